"""One GEMM shape, a few launches (for ncu): python scripts/prof_gemm.py M N K epi tile_n cta_pair"""
import sys
import torch
sys.path.insert(0, ".")
from tdm_b200 import ops
M, N, K = (int(x) for x in sys.argv[1:4])
epi, tile_n, pair = sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
a = torch.randn(M, K, device="cuda").bfloat16()
w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
bias = torch.randn(N, device="cuda")
o16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
o32 = torch.empty(M, N, device="cuda")
res = torch.randn(M, N, device="cuda")
gate = torch.randn(8, N, device="cuda")
kw = {"bias": dict(bias=bias, out_bf16=o16), "plain": dict(out_bf16=o16), "gelu": dict(bias=bias, act=ops.ACT_GELU_TANH, out_bf16=o16),
      "gate_res_2out": dict(bias=bias, gate=gate, rows_per_gate=M // 8, res=res, out_f32=o32, out_bf16=o16)}[epi]
for _ in range(4):
    ops.gemm(a, w, tile_n=tile_n, cta_pair=pair, **kw)
torch.cuda.synchronize()
print("ok")
