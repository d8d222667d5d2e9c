"""Two-stream attention forward (attention_fwd2.cu) against fp32 SDPA and against the one-tile kernel (TDM_B200_ATTN_V1=1):
accuracy on the PixArt shapes and the ragged test shapes, then timings of both kernels in one process."""
import json
import os
import sys
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
from _timing import timeit
from tdm_b200 import ops

torch.manual_seed(0)


def ref_attn(q, k, v, scale, bias):
    s = torch.einsum("bqhd,bkhd->bhqk", q.float(), k.float()) * scale
    if bias is not None:
        s = s + bias[:, None, None, :]
    return torch.einsum("bhqk,bkhd->bqhd", torch.softmax(s, -1), v.float()), torch.logsumexp(s, -1)


def check(B, H, Sq, Sk, d, cross, big_max=False):
    D = H * d
    q = torch.randn(B, Sq, H, d, device="cuda").bfloat16()
    k = torch.randn(B, Sk, H, d, device="cuda").bfloat16()
    v = torch.randn(B, Sk, H, d, device="cuda").bfloat16()
    if big_max:  # rows whose maximum grows by far more than 2^8 between tiles: the lazy-rescale path
        k[:, Sk // 2:] *= 6.0
    bias = None
    if cross:
        lens = torch.randint(1, Sk + 1, (B,), device="cuda")
        bias = (1 - (torch.arange(Sk, device="cuda")[None] < lens[:, None]).float()) * -10000.0
    o_ref, lse_ref = ref_attn(q, k, v, d ** -0.5, bias)
    out = {}
    for v1 in ("0", "1"):
        os.environ["TDM_B200_ATTN_V1"] = v1
        o = torch.full((B, Sq, H, d), float("nan"), device="cuda", dtype=torch.bfloat16)
        lse = torch.full((B, H, Sq), float("nan"), device="cuda")
        ops.attn_fwd(q, k, v, o, d ** -0.5, bias=bias, lse=lse)
        torch.cuda.synchronize()
        out["v1" if v1 == "1" else "v2"] = {
            "o_rel": ((o.float() - o_ref).abs().max() / o_ref.abs().max()).item(),
            "lse_abs": (lse - lse_ref).abs().max().item()}
    os.environ["TDM_B200_ATTN_V1"] = "0"
    print(json.dumps({"shape": [B, H, Sq, Sk, d], "cross": cross, "big_max": big_max, **out}), flush=True)
    return out["v2"]["o_rel"] < 1e-2 and out["v2"]["lse_abs"] < 2e-3


ok = True
for shp in [(1, 1, 256, 128, 64, False), (1, 1, 256, 256, 32, False), (2, 4, 256, 256, 64, False), (1, 2, 256, 256, 72, False),
            (2, 16, 1024, 1024, 72, False), (2, 16, 1024, 120, 72, True), (1, 2, 200, 77, 72, True), (1, 2, 384, 300, 40, True),
            (1, 3, 130, 250, 96, True), (8, 16, 1024, 1024, 72, False), (1, 16, 4096, 4096, 72, False)]:
    ok &= check(*shp)
ok &= check(2, 4, 512, 1024, 72, False, big_max=True)
print(json.dumps({"all_ok": bool(ok)}), flush=True)
if not ok or "--no-time" in sys.argv:
    sys.exit(0 if ok else 1)

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
for rep in range(2):
    for v1 in ("1", "0"):
        os.environ["TDM_B200_ATTN_V1"] = v1
        tag = "one-tile" if v1 == "1" else "two-stream"
        us, mhz = timeit(lambda i: ops.attn_fwd(q[i % NB], k[i % NB], v[i % NB], o[i % NB], d ** -0.5, lse=lse[i % NB]), iters=50)
        print(json.dumps({"kernel": tag, "case": "self 1024 B=8", "us": round(us, 2), "sm_mhz": mhz, "tflops_alg": round(4.0 * B * H * S * S * d / us / 1e6, 1)}), flush=True)
        us, mhz = timeit(lambda i: ops.attn_fwd(q[i % NB], k2[i % NB], v2[i % NB], o[i % NB], d ** -0.5, bias=bias, lse=lse[i % NB]), iters=50)
        print(json.dumps({"kernel": tag, "case": "cross 1024x120 B=8", "us": round(us, 2), "sm_mhz": mhz, "tflops_alg": round(4.0 * B * H * S * L * d / us / 1e6, 1)}), flush=True)
        us, mhz = timeit(lambda i: ops.attn_fwd(q4[i % 2], k4[i % 2], v4[i % 2], o4[i % 2], d ** -0.5), iters=20)
        print(json.dumps({"kernel": tag, "case": "self 4096 B=2", "us": round(us, 2), "sm_mhz": mhz, "tflops_alg": round(4.0 * 2 * H * S4 * S4 * d / us / 1e6, 1)}), flush=True)
os.environ["TDM_B200_ATTN_V1"] = "0"
os.makedirs("gpurun_out", exist_ok=True)
os.environ["TDM_B200_ATTN_TRACE"] = "gpurun_out/attn2_trace.txt"
ops.attn_fwd(q[0], k[0], v[0], o[0], d ** -0.5, lse=lse[0])
torch.cuda.synchronize()
os.environ["TDM_B200_ATTN_TRACE"] = ""
