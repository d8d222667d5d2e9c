#!/bin/bash
# Round-2 first visit: new parity tests, parity statistics, tcgen05-aware ncu metric names, a default bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
free -g > gpurun_out/host_mem.txt; nproc >> gpurun_out/host_mem.txt
timeout 900 python -m pytest -m gpu -q tests/test_golden_gpu.py tests/test_optimizer_gpu.py tests/test_export_checkpoint.py tests/test_iteration_gpu.py 2>&1 | tail -40 > gpurun_out/pytest_new.log
tail -15 gpurun_out/pytest_new.log
timeout 900 python scripts/parity_stats.py > gpurun_out/parity_stats.log 2>&1; echo "parity_stats rc=$?"
ncu --query-metrics 2>/dev/null | grep -iE 'tensor|umma|utc|tmem|tcgen' > gpurun_out/ncu_tensor_metrics.txt; wc -l gpurun_out/ncu_tensor_metrics.txt
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_v1.log 2> gpurun_out/bench_r2_v1.err; echo "bench rc=$?"
tail -c 4000 gpurun_out/bench_r2_v1.log; tail -5 gpurun_out/bench_r2_v1.err
