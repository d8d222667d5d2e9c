import sys, torch
sys.path.insert(0, ".")
from tdm_b200 import ops
M, D, N = 8192, 1152, 1024
x = [torch.randn(M, D, device="cuda") for _ in range(4)]
mod = torch.randn(8, 6, D, device="cuda")
for i in range(8):
    ops.ln_modulate_fwd(x[i % 4], mod[:, 0], mod[:, 1], N, 1e-6)
torch.cuda.synchronize(); print("ok")
