#!/bin/bash
# Per-launch device times of one distillation iteration (ncu, serialised, cold cache: compare SHARES).
# usage: gpu_launchlist.sh <skip> <count>
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --sample-batch 0"
$CMD > gpurun_out/plain.log 2> gpurun_out/plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s ${1:-17950} -c ${2:-6000} --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
echo "rc=$?"; tail -c 600 gpurun_out/plain.log; tail -3 gpurun_out/ncu.log; wc -l gpurun_out/launches.csv
