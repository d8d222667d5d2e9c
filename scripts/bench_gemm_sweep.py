"""GEMM cost model sweep: time = fixed + tiles * per_tile + K * rate, for our kernels and cuBLAS (torch.matmul).
One JSON line per case.  Rotating operands larger than L2."""
import json
import sys
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
from tdm_b200 import ops

DEV = "cuda"
NB = 4


from _timing import timeit as _timeit


def timeit(fn, iters=20, warm=3):
    return _timeit(fn, iters, warm)[0]


def case(M, N, K, variants):
    a = [torch.randn(M, K, device=DEV).bfloat16() for _ in range(NB)]
    w = [(torch.randn(N, K, device=DEV) * 0.05).bfloat16() for _ in range(NB)]
    o = [torch.empty(M, N, device=DEV, dtype=torch.bfloat16) for _ in range(NB)]
    row = {"M": M, "N": N, "K": K}
    for name, tn, pr in variants:
        if name == "cublas":
            us, mhz = _timeit(lambda i: torch.matmul(a[i % NB], w[i % NB].t(), out=o[i % NB]))
        else:
            us, mhz = _timeit(lambda i: ops.gemm(a[i % NB], w[i % NB], out_bf16=o[i % NB], tile_n=tn, cta_pair=pr))
        row[name] = round(us, 1)
        row[name + "_mhz"] = mhz
        row[name + "_tf"] = round(2.0 * M * N * K / us / 1e6)
    print(json.dumps(row), flush=True)


V = [("cublas", 0, 0), ("s192", 192, 1), ("s256", 256, 1), ("p192", 192, 2), ("p256", 256, 2), ("p128", 128, 2)]
if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "k"):
        for K in (128, 256, 576, 1152, 2304, 4608, 9216):
            case(8192, 4608, K, V)
    if which in ("all", "n"):
        for N in (1152, 2304, 3456, 4608, 9216):
            case(8192, N, 1152, V)
    if which in ("all", "m"):
        for M in (2048, 4096, 16384):
            case(M, 4608, 1152, V)
