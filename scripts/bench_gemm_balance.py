"""N = 1152 / 3456 GEMMs of a DiT block (M = 8192) with and without the narrow-tail MMA + balanced unit order."""
import json, os, sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "scripts")
from _timing import timeit
from tdm_b200 import ops
NB = 3
M = 8192
for name, N, K, kw in (("q2   bias->bf16", 1152, 1152, "plain"), ("o    bias+gate+res->f32+bf16", 1152, 1152, "o"), ("qkv  bias->bf16", 3456, 1152, "plain"),
                       ("ff2  bias+gate+res->f32", 1152, 4608, "ff2"), ("ff1  bias+gelu (no tail)", 4608, 1152, "plain"),
                       ("qkv dgrad (b mn)", 1152, 3456, "dgrad")):
    a = [torch.randn(M, K, device="cuda").bfloat16() for _ in range(NB)]
    w = [(torch.randn(K, N, device="cuda") * 0.05).bfloat16() if kw == "dgrad" else (torch.randn(N, K, device="cuda") * 0.05).bfloat16() for _ in range(NB)]
    bias = torch.randn(N, device="cuda")
    gate = torch.randn(8, N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    o16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    o32 = torch.empty(M, N, device="cuda")
    def run(i):
        if kw == "plain":
            ops.gemm(a[i % NB], w[i % NB], bias=bias, out_bf16=o16)
        elif kw == "dgrad":
            ops.gemm(a[i % NB], w[i % NB], b_mn_major=True, out_bf16=o16)
        elif kw == "o":
            ops.gemm(a[i % NB], w[i % NB], bias=bias, gate=gate, rows_per_gate=1024, res=res, out_f32=o32, out_bf16=o16)
        else:
            ops.gemm(a[i % NB], w[i % NB], bias=bias, gate=gate, rows_per_gate=1024, res=res, out_f32=o32)
    row = {"case": name, "M": M, "N": N, "K": K}
    for flag in ("0", "1", "0", "1"):
        os.environ["TDM_B200_GEMM_BALANCE"] = flag
        us, mhz = timeit(run, iters=60, min_ms=60.0)
        row.setdefault("balanced_us" if flag == "1" else "identity_us", []).append(round(us, 2))
    print(json.dumps(row), flush=True)
