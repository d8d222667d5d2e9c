"""Timing ablations of attn_fwd_kernel (TDM_B200_ATTN_DBG instantiations; outputs are wrong by construction) on the
PixArt-512 self-attention shape: which part of the per-tile softmax chain costs what."""
import json
import os
import sys
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
from _timing import timeit
from tdm_b200 import ops

B, H, S, d = 8, 16, 1024, 72
D = H * d
NB = 4
qkv = [torch.randn(B, S, 3 * D, device="cuda").bfloat16() for _ in range(NB)]
q = [t[:, :, :D].view(B, S, H, d) for t in qkv]
k = [t[:, :, D:2 * D].view(B, S, H, d) for t in qkv]
v = [t[:, :, 2 * D:].view(B, S, H, d) for t in qkv]
o = [torch.empty(B, S, H, d, device="cuda", dtype=torch.bfloat16) for _ in range(NB)]
lse = [torch.empty(B, H, S, device="cuda") for _ in range(NB)]
def case(tag, fn, flops, iters=50):
    us, mhz = timeit(fn, iters=iters)
    print(json.dumps({"case": tag, "us": round(us, 2), "sm_mhz": mhz, "tflops_alg": round(flops / us / 1e6, 1)}), flush=True)


names = {512: "no K/V TMA loads", 927: "handshakes only (no TMA, no MMA, no softmax)", 128: "no PV MMA", 256: "no QK MMA", 415: "no MMA, no softmax", 0: "baseline", 1: "no MUFU", 2: "no max exchange barrier", 4: "no P store", 8: "no S load", 16: "no max loop", 31: "skeleton (all off)"}
for dbg in (0, 1, 2, 4, 8, 16, 31, 128, 256, 415, 512, 927, 0):
    os.environ["TDM_B200_ATTN_DBG"] = str(dbg)
    us, mhz = timeit(lambda i: ops.attn_fwd(q[i % NB], k[i % NB], v[i % NB], o[i % NB], d ** -0.5, lse=lse[i % NB]), iters=50)
    tiles = B * H * (S // 128) ** 2
    print(json.dumps({"dbg": dbg, "what": names[dbg], "us": round(us, 2), "sm_mhz": mhz,
                      "clk_per_tile_per_sm": round(us * 1e-6 * (mhz or 1700) * 1e6 / (tiles / 148), 0),
                      "tflops_alg": round(4.0 * B * H * S * S * d / us / 1e6, 1)}), flush=True)
os.environ["TDM_B200_ATTN_DBG"] = "0"
# accuracy of the variants that are supposed to stay correct
ref = torch.nn.functional.scaled_dot_product_attention(q[0].transpose(1, 2).float(), k[0].transpose(1, 2).float(), v[0].transpose(1, 2).float()).transpose(1, 2)
for dbg in (0,):
    os.environ["TDM_B200_ATTN_DBG"] = str(dbg)
    ops.attn_fwd(q[0], k[0], v[0], o[0], d ** -0.5, lse=lse[0])
    torch.cuda.synchronize()
    err = (o[0].float() - ref).abs()
    print(json.dumps({"dbg": dbg, "max_abs_err": err.max().item(), "mean_abs_err": err.mean().item(), "ref_absmax": ref.abs().max().item()}), flush=True)
os.environ["TDM_B200_ATTN_DBG"] = "0"
# per-tile clock64 timeline of CTA 0 (one launch)
os.makedirs("gpurun_out", exist_ok=True)
os.environ["TDM_B200_ATTN_TRACE"] = "gpurun_out/attn_trace.txt"
ops.attn_fwd(q[0], k[0], v[0], o[0], d ** -0.5, lse=lse[0])
torch.cuda.synchronize()
os.environ["TDM_B200_ATTN_TRACE"] = ""
# cross-attention shape (one KV tile per item)
L = 120
kv = [torch.randn(B, L, 2 * D, device="cuda").bfloat16() for _ in range(NB)]
k2 = [t[:, :, :D].view(B, L, H, d) for t in kv]
v2 = [t[:, :, D:].view(B, L, H, d) for t in kv]
bias = torch.zeros(B, L, device="cuda")
us, mhz = timeit(lambda i: ops.attn_fwd(q[i % NB], k2[i % NB], v2[i % NB], o[i % NB], d ** -0.5, bias=bias, lse=lse[i % NB]), iters=50)
print(json.dumps({"case": "cross 1024x120", "us": round(us, 2), "sm_mhz": mhz, "tflops_alg": round(4.0 * B * H * S * L * d / us / 1e6, 1)}), flush=True)
# 4096-token self-attention (PixArt-1024), batch 2
S4 = 4096
qkv4 = [torch.randn(2, S4, 3 * D, device="cuda").bfloat16() for _ in range(2)]
q4 = [t[:, :, :D].view(2, S4, H, d) for t in qkv4]
k4 = [t[:, :, D:2 * D].view(2, S4, H, d) for t in qkv4]
v4 = [t[:, :, 2 * D:].view(2, S4, H, d) for t in qkv4]
o4 = [torch.empty(2, S4, H, d, device="cuda", dtype=torch.bfloat16) for _ in range(2)]
us, mhz = timeit(lambda i: ops.attn_fwd(q4[i % 2], k4[i % 2], v4[i % 2], o4[i % 2], d ** -0.5), iters=20)
print(json.dumps({"case": "self 4096 B=2", "us": round(us, 2), "sm_mhz": mhz, "tflops_alg": round(4.0 * 2 * H * S4 * S4 * d / us / 1e6, 1)}), flush=True)
