#!/bin/bash
# One GPU-box visit: the full GPU parity suite, smoke(), the reference arm and the default bench. Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
( time timeout 2400 python -m pytest tests -m gpu -q "$@" ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
( time timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/bench_reference.log 2> gpurun_out/bench_reference.err; echo "reference rc=$?"
tail -c 1500 gpurun_out/bench_reference.log; tail -4 gpurun_out/bench_reference.err
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.log; tail -6 gpurun_out/bench.err
