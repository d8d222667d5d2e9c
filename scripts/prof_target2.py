"""ncu target, round-2 kernels: attention forward / backward after the hand-off changes, the TAESD convolution, the
one-launch LoRA merge, the LN-modulate forward (HBM-bound).  A few launches each."""
import sys
import torch
sys.path.insert(0, ".")
from tdm_b200 import harness, ops
from tdm_b200.taesd import AutoencoderTiny

dev = "cuda"
B, H, S, d, L = 8, 16, 1024, 72, 120
D, M = H * d, 8 * 1024
torch.manual_seed(0)
qkv = torch.randn(B, S, 3 * D, device=dev).bfloat16()
q, k, v = (qkv[:, :, i * D:(i + 1) * D].view(B, S, H, d) for i in range(3))
o = torch.empty(B, S, H, d, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, S, device=dev)
do = torch.randn_like(o)
dq, dk, dv = (torch.empty(B, S, H, d, device=dev, dtype=torch.bfloat16) for _ in range(3))
x = torch.randn(M, D, device=dev)
mod = torch.randn(B, 6 * D, device=dev)
vae = AutoencoderTiny(device=dev, seed=0)
lat = torch.randn(8, 4, 64, 64, device=dev)
student, _, _ = harness.make_models("pixart-512", seed=0, device=dev, lora_rank=64, teacher=False, fake=False)
for _ in range(3):
    ops.attn_fwd(q, k, v, o, d ** -0.5, lse=lse)
    ops.attn_bwd(q, k, v, o, do, lse, d ** -0.5, dq, dk, dv)
    ops.ln_modulate_fwd(x, mod[:, :D], mod[:, D:2 * D], S)
    vae.decode(lat)
    student._eff_key = None
    student._refresh_merged()
torch.cuda.synchronize()
print("ok")
