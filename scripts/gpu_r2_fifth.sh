#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest -m gpu -x -q tests/test_taesd.py tests/test_fullsize_gpu.py tests/test_optimizer_gpu.py 2>&1 | tail -12 > gpurun_out/pytest_5.log; tail -6 gpurun_out/pytest_5.log
timeout 900 python bench.py --config pixart-1024 --steps 3 --warmup 3 --no-cpu-baseline --no-library-baseline --sample-batch 16 > gpurun_out/bench_r2_1024_1gpu.log 2> gpurun_out/bench_r2_1024_1gpu.err; echo "bench1024 rc=$?"; tail -c 2500 gpurun_out/bench_r2_1024_1gpu.log; tail -3 gpurun_out/bench_r2_1024_1gpu.err
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_r2_v3.log 2> gpurun_out/bench_r2_v3.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_r2_v3.log').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['step_roofline']['frac'], json.dumps(d['sampling']))"
