"""Wave-quantisation probe: time per row at M = 8192 (160 pair units on 74 pairs: 3 rounds) vs M = 7424 (145 units: 2 rounds)."""
import json, sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "scripts")
from tdm_b200 import ops
from _timing import timeit
NB = 4
for (N, K) in ((1152, 1152), (1152, 4608), (3456, 1152)):
    for M in (7424, 8192):
        a = [torch.randn(M, K, device="cuda").bfloat16() for _ in range(NB)]
        w = [(torch.randn(N, K, device="cuda") * 0.05).bfloat16() for _ in range(NB)]
        o = [torch.empty(M, N, device="cuda", dtype=torch.bfloat16) for _ in range(NB)]
        row = {"M": M, "N": N, "K": K}
        for name, tn, pr in (("p256", 256, 2), ("p192", 192, 2), ("cublas", 0, 0), ("p256b", 256, 2)):
            if name == "cublas":
                us, mhz = timeit(lambda i: torch.matmul(a[i % NB], w[i % NB].t(), out=o[i % NB]))
            else:
                us, mhz = timeit(lambda i: ops.gemm(a[i % NB], w[i % NB], out_bf16=o[i % NB], tile_n=tn, cta_pair=pr))
            row[name] = [round(us, 1), mhz, round(us / M * 8192, 1)]
        print(json.dumps(row), flush=True)
