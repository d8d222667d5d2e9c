"""Backward GEMMs of a DiT block (tokens = 8192) against the tile configuration, device time from CUDA-graph replays."""
import json, sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "scripts")
from _timing import timeit_graph
from tdm_b200 import ops
NB = 3
T = 8192
cfgs = ((0, 0), (256, 2), (192, 2), (128, 2), (256, 1), (192, 1), (128, 1))
def tag(tn, cp): return "auto" if tn == 0 else str(tn) + ("p" if cp == 2 else "s")
# dgrads: dx[T, k_in] = dy[T, n_out] . W[n_out, k_in]  (B MN-major)
for name, n_out, k_in in (("ff2 dgrad (gelu' in the model)", 1152, 4608), ("o / q2 / o2 dgrad", 1152, 1152), ("kv-less qkv dgrad", 3456, 1152), ("ff1 dgrad", 4608, 1152)):
    dy = [torch.randn(T, n_out, device="cuda").bfloat16() for _ in range(NB)]
    w = [(torch.randn(n_out, k_in, device="cuda") * 0.05).bfloat16() for _ in range(NB)]
    o16 = torch.empty(T, k_in, device="cuda", dtype=torch.bfloat16)
    row = {"case": name, "N": k_in, "K": n_out}
    for tn, cp in cfgs:
        if cp == 2 and tn == 192: continue
        try:
            us, _ = timeit_graph(lambda i: ops.gemm(dy[i % NB], w[i % NB], b_mn_major=True, out_bf16=o16, tile_n=tn, cta_pair=cp), n=40)
            row[tag(tn, cp)] = round(us, 1)
        except Exception:  # noqa: BLE001
            row[tag(tn, cp)] = "err"
    print(json.dumps(row), flush=True)
# wgrads: dW[n_out, k_in] += dy^T x
for name, n_out, k_in in (("w_o", 1152, 1152), ("w_qkv", 3456, 1152), ("w_ff1", 4608, 1152), ("w_ff2", 1152, 4608)):
    dy = [torch.randn(T, n_out, device="cuda").bfloat16() for _ in range(NB)]
    x = [torch.randn(T, k_in, device="cuda").bfloat16() for _ in range(NB)]
    dw = torch.zeros(n_out, k_in, device="cuda")
    row = {"case": name, "M": n_out, "N": k_in}
    for tn, cp in cfgs:
        if cp == 2 and tn == 192: continue
        try:
            us, _ = timeit_graph(lambda i: ops.gemm(dy[i % NB], x[i % NB], a_mn_major=True, b_mn_major=True, res=dw, out_f32=dw, tile_n=tn, cta_pair=cp), n=30)
            row[tag(tn, cp)] = round(us, 1)
        except Exception:  # noqa: BLE001
            row[tag(tn, cp)] = "err"
    print(json.dumps(row), flush=True)
