"""Backward GEMM shapes (dgrad: B MN-major; wgrad: both MN-major, fp32 accumulate): auto choice vs forced variants."""
import json, sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "scripts")
from tdm_b200 import ops
from _timing import timeit
NB = 4
def dgrad(M, N, K):
    dy = [torch.randn(M, K, device="cuda").bfloat16() for _ in range(NB)]
    w = [(torch.randn(K, N, device="cuda") * 0.05).bfloat16() for _ in range(NB)]
    o = [torch.empty(M, N, device="cuda", dtype=torch.bfloat16) for _ in range(NB)]
    row = {"kind": "dgrad", "M": M, "N": N, "K": K}
    for name, tn, pr in (("auto", 0, 0), ("s192", 192, 1), ("s256", 256, 1), ("p256", 256, 2), ("p128", 128, 2), ("auto2", 0, 0)):
        us, mhz = timeit(lambda i: ops.gemm(dy[i % NB], w[i % NB], b_mn_major=True, out_bf16=o[i % NB], tile_n=tn, cta_pair=pr))
        row[name] = [round(us, 1), mhz]
    print(json.dumps(row), flush=True)
def wgrad(tokens, n_out, k_in):
    dy = [torch.randn(tokens, n_out, device="cuda").bfloat16() for _ in range(NB)]
    x = [torch.randn(tokens, k_in, device="cuda").bfloat16() for _ in range(NB)]
    dw = torch.zeros(n_out, k_in, device="cuda")
    row = {"kind": "wgrad", "tokens": tokens, "n_out": n_out, "k_in": k_in}
    for name, tn, pr, sk in (("auto", 0, 0, 0), ("s192", 192, 1, 0), ("s192_1", 192, 1, 1), ("p256", 256, 2, 0), ("p256_1", 256, 2, 1), ("p128", 128, 2, 0), ("auto2", 0, 0, 0)):
        try:
            us, mhz = timeit(lambda i: ops.gemm(dy[i % NB], x[i % NB], a_mn_major=True, b_mn_major=True, res=dw, out_f32=dw, tile_n=tn, cta_pair=pr, split_k=sk))
            row[name] = [round(us, 1), mhz]
        except Exception as e:
            row[name] = str(e)[-60:]
    print(json.dumps(row), flush=True)
dgrad(8192, 1152, 1152); dgrad(8192, 1152, 3456); dgrad(8192, 4608, 1152); dgrad(8192, 1152, 4608)
wgrad(8192, 1152, 1152); wgrad(8192, 3456, 1152); wgrad(8192, 4608, 1152); wgrad(8192, 1152, 4608); wgrad(8192, 1152, 64); wgrad(8192, 64, 1152); wgrad(8192, 192, 1152)
