#!/bin/bash
# Run on the GPU box through gpurun: GPU parity tests (all of tests/ unless files are given); logs to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
if [ $# -eq 0 ]; then set -- tests; fi
python -m pytest -m gpu -q "$@" 2>&1 | tail -80 | tee gpurun_out/pytest_gpu.log
