"""Timings of the two-stream attention forward: self 1024 (B=8), cross 1024x120, self 4096 (B=2); optional ablation
variants (TDM_B200_ATTN2_DBG) and the clock64 trace of CTA 0."""
import json
import os
import sys
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
from _timing import timeit
from tdm_b200 import ops

B, H, S, d, L = 8, 16, 1024, 72, 120
D = H * d
NB = 4
qkv = [torch.randn(B, S, 3 * D, device="cuda").bfloat16() for _ in range(NB)]
q = [t[:, :, :D].view(B, S, H, d) for t in qkv]
k = [t[:, :, D:2 * D].view(B, S, H, d) for t in qkv]
v = [t[:, :, 2 * D:].view(B, S, H, d) for t in qkv]
o = [torch.empty(B, S, H, d, device="cuda", dtype=torch.bfloat16) for _ in range(NB)]
lse = [torch.empty(B, H, S, device="cuda") for _ in range(NB)]
kv = [torch.randn(B, L, 2 * D, device="cuda").bfloat16() for _ in range(NB)]
k2 = [t[:, :, :D].view(B, L, H, d) for t in kv]
v2 = [t[:, :, D:].view(B, L, H, d) for t in kv]
bias = torch.zeros(B, L, device="cuda")
S4 = 4096
qkv4 = [torch.randn(2, S4, 3 * D, device="cuda").bfloat16() for _ in range(2)]
q4 = [t[:, :, :D].view(2, S4, H, d) for t in qkv4]
k4 = [t[:, :, D:2 * D].view(2, S4, H, d) for t in qkv4]
v4 = [t[:, :, 2 * D:].view(2, S4, H, d) for t in qkv4]
o4 = [torch.empty(2, S4, H, d, device="cuda", dtype=torch.bfloat16) for _ in range(2)]
variants = [int(x) for x in sys.argv[1:]] or [0]
for dbg in variants:
    os.environ["TDM_B200_ATTN2_DBG"] = str(dbg)
    us, mhz = timeit(lambda i: ops.attn_fwd(q[i % NB], k[i % NB], v[i % NB], o[i % NB], d ** -0.5, lse=lse[i % NB]), iters=50)
    print(json.dumps({"dbg": dbg, "case": "self 1024 B=8", "us": round(us, 2), "sm_mhz": mhz, "tflops_alg": round(4.0 * B * H * S * S * d / us / 1e6, 1)}), flush=True)
    us, mhz = timeit(lambda i: ops.attn_fwd(q[i % NB], k2[i % NB], v2[i % NB], o[i % NB], d ** -0.5, bias=bias, lse=lse[i % NB]), iters=50)
    print(json.dumps({"dbg": dbg, "case": "cross 1024x120 B=8", "us": round(us, 2), "sm_mhz": mhz, "tflops_alg": round(4.0 * B * H * S * L * d / us / 1e6, 1)}), flush=True)
    us, mhz = timeit(lambda i: ops.attn_fwd(q4[i % 2], k4[i % 2], v4[i % 2], o4[i % 2], d ** -0.5), iters=20)
    print(json.dumps({"dbg": dbg, "case": "self 4096 B=2", "us": round(us, 2), "sm_mhz": mhz, "tflops_alg": round(4.0 * 2 * H * S4 * S4 * d / us / 1e6, 1)}), flush=True)
os.environ["TDM_B200_ATTN2_DBG"] = "0"
os.makedirs("gpurun_out", exist_ok=True)
for name, kk, vv, bb in (("self", k, v, None), ("cross", k2, v2, bias)):
    os.environ["TDM_B200_ATTN_TRACE"] = f"gpurun_out/attn2_trace_{name}.txt"
    ops.attn_fwd(q[0], kk[0], vv[0], o[0], d ** -0.5, bias=bb, lse=lse[0])
    torch.cuda.synchronize()
os.environ["TDM_B200_ATTN_TRACE"] = ""
