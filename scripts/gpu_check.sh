#!/bin/bash
# Run on the GPU box through gpurun: GPU parity tests, then optional extra commands; logs to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests -m gpu -x -q "$@" 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
