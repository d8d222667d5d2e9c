"""From an `ncu --set full` capture of scripts/prof_target.py (read here with `ncu -i X.ncu-rep --page raw --csv > raw.csv`):
profiles/r02_ncu_tensor_pipe_summary.txt (tcgen05-aware tensor-pipe counters per kernel) and profiles/r02_gemm_traffic.json
(DRAM bytes per launch of the dominant GEMM, which bench.py reports as roofline.traffic).
    python scripts/ncu_traffic.py raw.csv"""
import csv
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
WANT = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum', 'sm__sass_inst_executed_op_utcmma.sum',
        'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__mem_tensor_reads_op_ldt.sum', 'sm__mem_tensor_writes_op_stt.sum',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active']
NAMES = ['gemm2 FF1 8192x4608x1152 bias+GELU->bf16', 'gemm2 o-proj 8192x1152x1152 bias+gate+res->f32+bf16',
         'attn_fwd self B8 H16 S1024 d72', 'attn_fwd cross Sk=120 (bias)', 'attn_fwd self (repeat)', 'attn_bwd self']


def main(path, tag="r02"):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = ["# ncu --set full --clock-control none (+ tcgen05-aware tensor counters), scripts/prof_target.py, B200.",
           "# The tensor-pipe utilisation metric that SEES tcgen05.mma (UTCHMMA) is sm__pipe_tensor_subpipe_hmma_cycles_active",
           "# (peak = 8192 dense bf16 FLOP/clk/SM); sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum equals the executed 2MNK",
           "# exactly (8.697e10 for FF1; 4.295e10 for self-attention = 4*B*H*S*S*80, head_dim padded 72 -> 80).",
           "# (round 1's sm__pipe_tensor_cycles_active_realtime = 3 % was the legacy-HMMA 'realtime' counter; it does not count UTCHMMA.)"]
    for r, nm in zip(rows[2:], NAMES):
        out.append("== " + nm)
        out.append("kernel = " + r[idx['Kernel Name']][:100])
        for w in WANT:
            if w in idx:
                out.append(f"{w} = {r[idx[w]]} {units[idx[w]]}")
        fl, us = float(r[idx['sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum']]), float(r[idx['gpu__time_duration.sum']])
        out.append(f"executed tensor TFLOP/s (under ncu) = {fl / us / 1e6:.1f}")
    (ROOT / "profiles" / f"{tag}_ncu_tensor_pipe_summary.txt").write_text("\n".join(out) + "\n")
    r = rows[2]
    rd, wr = float(r[idx['dram__bytes_read.sum']]), float(r[idx['dram__bytes_write.sum']])
    tr = {"bytes_per_launch": (rd + wr) * 1e6, "read_MB": rd, "write_MB": wr, "shape": "8192x4608x1152 (FF1, bias + GELU-tanh -> bf16)",
          "algorithmic_bytes": 2 * (8192 * 1152 + 4608 * 1152 + 8192 * 4608),
          "source": "ncu --set full --clock-control none capture of scripts/prof_target.py -> profiles/%s_ncu_tensor_pipe_summary.txt (scripts/ncu_traffic.py)" % tag}
    (ROOT / "profiles" / f"{tag}_gemm_traffic.json").write_text(json.dumps(tr, indent=1))
    print("\n".join(out))


if __name__ == "__main__":
    main(*sys.argv[1:])
