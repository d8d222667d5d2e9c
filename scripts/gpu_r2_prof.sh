#!/bin/bash
# ncu counters (tcgen05-aware tensor pipe, DRAM) of the round-2 kernels, one family per invocation; then nothing else.
mkdir -p gpurun_out
M="gpu__time_duration.sum,sm__cycles_elapsed.avg,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum,sm__sass_inst_executed_op_utcmma.sum,sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size"
timeout 300 python scripts/prof_target2.py > gpurun_out/prof2_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof2_plain.log; exit 1; }
rm -f gpurun_out/r02_prof_v2_*.csv
for spec in "attn_fwd2_kernel 2 1" "attn_bwd_kernel 2 1" "ln_modulate_fwd 2 1" "lora_merge 2 1" "taesd_conv64 95 4"; do
  set -- $spec
  timeout 600 ncu --metrics $M --clock-control none -k regex:$1 -s $2 -c $3 --csv --log-file gpurun_out/r02_prof_v2_$1.csv python scripts/prof_target2.py > gpurun_out/ncu_v2_$1.log 2>&1
  echo "$1 rc=$? lines=$(wc -l < gpurun_out/r02_prof_v2_$1.csv)"
done
