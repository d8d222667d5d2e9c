"""ncu target: the two-stream attention forward and the backward on the PixArt-512 self-attention shape."""
import sys, torch
sys.path.insert(0, ".")
from tdm_b200 import ops
B, H, S, d = 8, 16, 1024, 72
D = H * d
torch.manual_seed(0)
qkv = torch.randn(B, S, 3 * D, device="cuda").bfloat16()
q, k, v = (qkv[:, :, i * D:(i + 1) * D].view(B, S, H, d) for i in range(3))
o = torch.empty(B, S, H, d, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(B, H, S, device="cuda")
do = torch.randn_like(o)
dqkv = torch.zeros(B, S, 3 * D, device="cuda", dtype=torch.bfloat16)
dq, dk, dv = (dqkv[:, :, i * D:(i + 1) * D].view(B, S, H, d) for i in range(3))
for _ in range(3):
    ops.attn_fwd(q, k, v, o, d ** -0.5, lse=lse)
    ops.attn_bwd(q, k, v, o, do, lse, d ** -0.5, dq, dk, dv)
torch.cuda.synchronize()
print("ok")
