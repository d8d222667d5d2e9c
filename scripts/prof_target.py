"""ncu target: a handful of launches of the three tensor-core kernel families at the PixArt-512 batch-8 shapes
(gemm2 FF1 bias+GELU, gemm2 o-proj gate+residual+2 outputs, attention forward self / cross, attention backward self)."""
import sys
import torch
sys.path.insert(0, ".")
from tdm_b200 import ops

dev = "cuda"
B, H, S, d, L = 8, 16, 1024, 72, 120
D, M = H * d, 8 * 1024
torch.manual_seed(0)
h = torch.randn(M, D, device=dev).bfloat16()
w1 = (torch.randn(4 * D, D, device=dev) * 0.03).bfloat16()
wo = (torch.randn(D, D, device=dev) * 0.03).bfloat16()
b1, bo = torch.randn(4 * D, device=dev), torch.randn(D, device=dev)
g = torch.empty(M, 4 * D, device=dev, dtype=torch.bfloat16)
x, x1 = torch.randn(M, D, device=dev), torch.empty(M, D, device=dev)
x1b = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
mod = torch.randn(B, 6 * D, device=dev)
qkv = torch.randn(B, S, 3 * D, device=dev).bfloat16()
q, k, v = (qkv[:, :, i * D:(i + 1) * D].view(B, S, H, d) for i in range(3))
o = torch.empty(B, S, H, d, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, S, device=dev)
kv = torch.randn(B, L, 2 * D, device=dev).bfloat16()
k2, v2 = kv[:, :, :D].view(B, L, H, d), kv[:, :, D:].view(B, L, H, d)
bias = torch.zeros(B, L, device=dev)
do = torch.randn_like(o)
dq, dk, dv = (torch.empty(B, S, H, d, device=dev, dtype=torch.bfloat16) for _ in range(3))
for _ in range(3):
    ops.gemm(h, w1, bias=b1, act=ops.ACT_GELU_TANH, out_bf16=g)
    ops.gemm(h, wo, bias=bo, gate=mod[:, 2 * D:3 * D], rows_per_gate=S, res=x, out_f32=x1, out_bf16=x1b)
    ops.attn_fwd(q, k, v, o, d ** -0.5, lse=lse)
    ops.attn_fwd(q, k2, v2, o, d ** -0.5, bias=bias, lse=lse)
    ops.attn_fwd(q, k, v, o, d ** -0.5, lse=lse)
    ops.attn_bwd(q, k, v, o, do, lse, d ** -0.5, dq, dk, dv)
torch.cuda.synchronize()
print("ok")
