#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/attn_trace.txt
timeout 300 python -m pytest -m gpu -x -q tests/test_block_kernels_gpu.py -k "attention" 2>&1 | tail -8 > gpurun_out/pytest_attn.log; tail -4 gpurun_out/pytest_attn.log
timeout 300 python scripts/attn_ablate.py > gpurun_out/attn_ablate2.jsonl 2> gpurun_out/attn_ablate2.err; echo "ablate rc=$?"; cat gpurun_out/attn_ablate2.jsonl; tail -3 gpurun_out/attn_ablate2.err
head -40 gpurun_out/attn_trace.txt
