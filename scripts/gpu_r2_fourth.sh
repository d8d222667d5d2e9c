#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest -m gpu -x -q tests/test_block_kernels_gpu.py tests/test_dit_gpu.py tests/test_iteration_gpu.py tests/test_golden_gpu.py 2>&1 | tail -12 > gpurun_out/pytest_4.log; tail -6 gpurun_out/pytest_4.log
timeout 300 python scripts/bench_kernels.py attn > gpurun_out/bench_kernels_attn_r2.log 2>&1; grep attn gpurun_out/bench_kernels_attn_r2.log | cut -c1-220
timeout 900 python scripts/parity_stats.py > gpurun_out/parity_stats.log 2>&1; echo "parity rc=$?"
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_r2_v2.log 2> gpurun_out/bench_r2_v2.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_r2_v2.log').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['step_roofline']['frac'], d['sampling']['value'], d['roofline']['frac'], d['clocks'])"
