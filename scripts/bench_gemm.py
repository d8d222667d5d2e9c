"""Micro-benchmark of tdm_gemm_bf16 on the DiT-XL/2 shapes (CUDA events, L2 flushed by rotating buffers)."""
import json
import sys
import torch
sys.path.insert(0, ".")
from tdm_b200 import ops

def bench(M, N, K, tile_n=0, iters=20, **kw):
    nbuf = 6
    a = [torch.randn(M, K, device="cuda").to(torch.bfloat16) for _ in range(nbuf)]
    b = [(torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16) for _ in range(nbuf)]
    o = [torch.empty(M, N, device="cuda", dtype=torch.bfloat16) for _ in range(nbuf)]
    for i in range(3):
        ops.gemm(a[i % nbuf], b[i % nbuf], out_bf16=o[i % nbuf], tile_n=tile_n, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        ops.gemm(a[i % nbuf], b[i % nbuf], out_bf16=o[i % nbuf], tile_n=tile_n, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    # cuBLAS for comparison (library baseline)
    for i in range(3):
        torch.matmul(a[i % nbuf], b[i % nbuf].t())
    torch.cuda.synchronize()
    e0.record()
    for i in range(iters):
        torch.matmul(a[i % nbuf], b[i % nbuf].t())
    e1.record()
    torch.cuda.synchronize()
    ms_ref = e0.elapsed_time(e1) / iters
    fl = 2.0 * M * N * K
    return {"M": M, "N": N, "K": K, "tile_n": tile_n, "us": ms * 1e3, "tflops": fl / ms / 1e9,
            "cublas_us": ms_ref * 1e3, "cublas_tflops": fl / ms_ref / 1e9}

if __name__ == "__main__":
    res = []
    for (M, N, K) in [(8192, 1152, 1152), (8192, 3456, 1152), (8192, 4608, 1152), (8192, 1152, 4608),
                      (16384, 4608, 1152), (960, 2304, 1152), (8192, 8192, 8192)]:
        for tn in (128, 256):
            r = bench(M, N, K, tn)
            print(json.dumps(r), flush=True)
            res.append(r)
