mkdir -p gpurun_out
timeout 200 python scripts/bench_kernels.py attn > gpurun_out/r02_kernel_microbench_attn_v2.jsonl 2>&1
TDM_B200_ATTN_V1=1 timeout 200 python scripts/bench_kernels.py attn 2>&1 | grep attn_fwd | sed 's/"kernel": "attn_fwd"/"kernel": "attn_fwd (one-tile kernel, TDM_B200_ATTN_V1=1)"/' >> gpurun_out/r02_kernel_microbench_attn_v2.jsonl
timeout 300 python scripts/bench_wgrad_splits.py > gpurun_out/r02_wgrad_split_sweep.jsonl 2>&1
timeout 300 python scripts/bench_gemm_balance.py > gpurun_out/r02_gemm_balance.jsonl 2>&1
timeout 100 python scripts/merge_time.py > gpurun_out/r02_lora_merge_time.json 2>&1
tail -3 gpurun_out/r02_kernel_microbench_attn_v2.jsonl; tail -1 gpurun_out/r02_lora_merge_time.json
