#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest -m gpu -q tests/test_optimizer_gpu.py tests/test_export_checkpoint.py 2>&1 | tail -30 > gpurun_out/pytest_new2.log; tail -5 gpurun_out/pytest_new2.log
timeout 300 python scripts/attn_ablate.py > gpurun_out/attn_ablate.jsonl 2> gpurun_out/attn_ablate.err; echo "ablate rc=$?"; cat gpurun_out/attn_ablate.jsonl
M="gpu__time_duration.sum,sm__cycles_elapsed.avg,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor_subpipe_hmma.sum,sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum,sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.avg.pct_of_peak_sustained_elapsed,sm__ops_path_tensor_src_bf16_dst_fp32.avg.pct_of_peak_sustained_elapsed,sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__mem_tensor_reads_op_ldt.sum,sm__mem_tensor_writes_op_stt.sum,sm__sass_inst_executed_op_utcmma.sum,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed"
timeout 200 python scripts/prof_target.py > gpurun_out/prof_plain.log 2>&1 && \
timeout 900 ncu --set full --metrics $M --clock-control none --import-source on -k regex:'gemm2_bf16|attn_fwd|attn_bwd_kernel' -s 12 -c 6 -o gpurun_out/r02_prof_v1 -f python scripts/prof_target.py > gpurun_out/ncu_v1.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_v1.log; ls -la gpurun_out/*.ncu-rep
