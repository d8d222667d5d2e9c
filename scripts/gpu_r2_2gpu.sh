#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest -m gpu -x -q tests/test_dit_gpu.py tests/test_iteration_gpu.py tests/test_dp_nccl_gpu.py 2>&1 | tail -15 > gpurun_out/pytest_nccl.log; tail -8 gpurun_out/pytest_nccl.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dp_timeline.py gpurun_out/dp_timeline_2gpu.json > gpurun_out/dp_timeline.log 2>&1; echo "timeline rc=$?"; tail -1 gpurun_out/dp_timeline.log | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --sample-batch 0 > gpurun_out/bench_r2_2gpu.log 2> gpurun_out/bench_r2_2gpu.err; echo "bench2 rc=$?"; python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_r2_2gpu.log') if l.startswith('{')][-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])"
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --sample-batch 0 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_r2_1gpu_ref.log 2>/dev/null; python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_r2_1gpu_ref.log') if l.startswith('{')][-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['gpu_launches'])"
