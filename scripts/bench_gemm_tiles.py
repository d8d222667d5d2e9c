"""Forward GEMMs of a DiT block (M = 8192) against the tile configuration (tile_n x cta_pair), device time from CUDA-graph
replays: checks the chooser of tdm_gemm_bf16 (tile_n = 0, cta_pair = 0 = automatic)."""
import json, sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "scripts")
from _timing import timeit_graph
from tdm_b200 import ops
NB = 3
M = 8192
for name, N, K, kw in (("q2   bias->bf16", 1152, 1152, "plain"), ("o    bias+gate+res->f32+bf16", 1152, 1152, "o"), ("o2   bias+res->f32", 1152, 1152, "o2"),
                       ("qkv  bias->bf16", 3456, 1152, "plain"), ("ff2  bias+gate+res->f32", 1152, 4608, "ff2"), ("ff1  bias+gelu->bf16", 4608, 1152, "ff1"),
                       ("ff1 dgrad (b mn) -> bf16", 1152, 4608, "dgrad"), ("qkv dgrad (b mn) -> bf16", 1152, 3456, "dgrad")):
    a = [torch.randn(M, K, device="cuda").bfloat16() for _ in range(NB)]
    w = [(torch.randn(K, N, device="cuda") * 0.05).bfloat16() if kw == "dgrad" else (torch.randn(N, K, device="cuda") * 0.05).bfloat16() for _ in range(NB)]
    bias = torch.randn(N, device="cuda"); gate = torch.randn(8, N, device="cuda"); res = torch.randn(M, N, device="cuda")
    o16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16); o32 = torch.empty(M, N, device="cuda")
    def run(i, tn, cp):
        kwargs = dict(tile_n=tn, cta_pair=cp)
        if kw == "plain": ops.gemm(a[i % NB], w[i % NB], bias=bias, out_bf16=o16, **kwargs)
        elif kw == "ff1": ops.gemm(a[i % NB], w[i % NB], bias=bias, act=ops.ACT_GELU_TANH, out_bf16=o16, **kwargs)
        elif kw == "dgrad": ops.gemm(a[i % NB], w[i % NB], b_mn_major=True, out_bf16=o16, **kwargs)
        elif kw == "o": ops.gemm(a[i % NB], w[i % NB], bias=bias, gate=gate, rows_per_gate=1024, res=res, out_f32=o32, out_bf16=o16, **kwargs)
        elif kw == "o2": ops.gemm(a[i % NB], w[i % NB], bias=bias, res=res, out_f32=o32, **kwargs)
        else: ops.gemm(a[i % NB], w[i % NB], bias=bias, gate=gate, rows_per_gate=1024, res=res, out_f32=o32, **kwargs)
    row = {"case": name, "N": N, "K": K}
    for tn, cp in ((0, 0), (256, 2), (192, 2), (128, 2), (256, 1), (192, 1), (128, 1)):
        if kw == "dgrad" and cp == 2 and tn == 192: continue
        try:
            us, _ = timeit_graph(lambda i: run(i, tn, cp), n=40)
            row[f"{'auto' if tn == 0 else str(tn) + ('p' if cp == 2 else 's')}"] = round(us, 1)
        except Exception as e:  # noqa: BLE001
            row[f"{tn}{'p' if cp == 2 else 's'}"] = "err"
    print(json.dumps(row), flush=True)
