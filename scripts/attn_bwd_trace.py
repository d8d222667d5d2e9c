"""clock64 trace of one attention-backward CTA (self-attention, PixArt-512 shape) + timing."""
import json, os, sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "scripts")
from _timing import timeit
from tdm_b200 import ops
B, H, S, d = 8, 16, 1024, 72
D = H * d
qkv = torch.randn(B, S, 3 * D, device="cuda").bfloat16()
q, k, v = (qkv[:, :, i * D:(i + 1) * D].view(B, S, H, d) for i in range(3))
o = torch.empty(B, S, H, d, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(B, H, S, device="cuda")
do = torch.randn_like(o)
dqkv = torch.zeros(B, S, 3 * D, device="cuda", dtype=torch.bfloat16)
dq, dk, dv = (dqkv[:, :, i * D:(i + 1) * D].view(B, S, H, d) for i in range(3))
ops.attn_fwd(q, k, v, o, d ** -0.5, lse=lse)
us, mhz = timeit(lambda i: ops.attn_bwd(q, k, v, o, do, lse, d ** -0.5, dq, dk, dv), iters=30)
print(json.dumps({"case": "bwd self 1024 B=8 (delta + main + cast)", "us": round(us, 2), "sm_mhz": mhz}), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
os.environ["TDM_B200_ATTN_TRACE"] = "gpurun_out/attn_bwd_trace.txt"
ops.attn_bwd(q, k, v, o, do, lse, d ** -0.5, dq, dk, dv)
torch.cuda.synchronize()
os.environ["TDM_B200_ATTN_TRACE"] = ""
