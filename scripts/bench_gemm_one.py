"""A few GEMM shapes, interleaved variants (tile_n, cta_pair) + cuBLAS, with clocks: python scripts/bench_gemm_one.py"""
import json, sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "scripts")
from tdm_b200 import ops
from _timing import timeit
NB = 4
for (M, N, K) in ((8192, 4608, 128), (8192, 4608, 1152), (8192, 1152, 1152), (8192, 3456, 1152)):
    a = [torch.randn(M, K, device="cuda").bfloat16() for _ in range(NB)]
    w = [(torch.randn(N, K, device="cuda") * 0.05).bfloat16() for _ in range(NB)]
    o = [torch.empty(M, N, device="cuda", dtype=torch.bfloat16) for _ in range(NB)]
    row = {"M": M, "N": N, "K": K}
    for name, tn, pr in (("s192", 192, 1), ("p256", 256, 2), ("p192", 192, 2), ("cublas", 0, 0), ("s192b", 192, 1), ("p256b", 256, 2)):
        if name == "cublas":
            us, mhz = timeit(lambda i: torch.matmul(a[i % NB], w[i % NB].t(), out=o[i % NB]))
        else:
            us, mhz = timeit(lambda i: ops.gemm(a[i % NB], w[i % NB], out_bf16=o[i % NB], tile_n=tn, cta_pair=pr))
        row[name] = [round(us, 1), mhz]
    print(json.dumps(row), flush=True)
