#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_r2_${N}gpu.log 2> gpurun_out/bench_r2_${N}gpu.err; echo "bench$N rc=$?"; python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_r2_${N}gpu.log') if l.startswith('{')][-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['sampling']['value'], d['sampling']['rgb']['value'], d['clocks'])"
tail -3 gpurun_out/bench_r2_${N}gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --config pixart-1024 --steps 4 --warmup 3 --sample-batch 16 > gpurun_out/bench_r2_1024_${N}gpu.log 2> gpurun_out/bench_r2_1024_${N}gpu.err; echo "bench1024_$N rc=$?"; python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_r2_1024_${N}gpu.log') if l.startswith('{')][-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['step_roofline']['frac'], d['sampling']['value'])"
tail -3 gpurun_out/bench_r2_1024_${N}gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/dp_timeline.py gpurun_out/dp_timeline_${N}gpu.json > gpurun_out/dp_timeline_$N.log 2>&1; tail -1 gpurun_out/dp_timeline_$N.log | cut -c1-500
