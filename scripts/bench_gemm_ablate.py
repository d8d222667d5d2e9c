import json, sys, torch
sys.path.insert(0, ".")
from tdm_b200 import ops
def t(M, N, K, tn, **kw):
    a = [torch.randn(M, K, device="cuda").bfloat16() for _ in range(4)]
    w = [(torch.randn(N, K, device="cuda") * 0.05).bfloat16() for _ in range(4)]
    o = [torch.empty(M, N, device="cuda", dtype=torch.bfloat16) for _ in range(4)]
    bias = torch.randn(N, device="cuda")
    f = lambda i: ops.gemm(a[i % 4], w[i % 4], out_bf16=o[i % 4], bias=bias, tile_n=tn, cta_pair=1)
    for i in range(3): f(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20): f(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20 * 1e3
for (M, N, K, tn) in ((8192, 3456, 1152, 192), (8192, 1152, 1152, 192), (8192, 4608, 1152, 256), (8192, 1152, 4608, 192)):
    print(json.dumps({"M": M, "N": N, "K": K, "tile_n": tn, "us": round(t(M, N, K, tn), 1)}), flush=True)
