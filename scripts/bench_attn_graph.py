"""Attention kernels, device time from CUDA-graph replays (the Python call of ops.attn_* costs 20-30 us, which bounds a plain
timing loop for the cross-attention shapes) next to the plain loop."""
import json, os, sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "scripts")
from _timing import timeit, timeit_graph
from tdm_b200 import ops
B, H, S, L, d = 8, 16, 1024, 120, 72
D = H * d
NB = 3
def make(cross):
    if not cross:
        qkv = torch.randn(B, S, 3 * D, device="cuda").bfloat16()
        q, k, v = (qkv[:, :, i * D:(i + 1) * D].view(B, S, H, d) for i in range(3))
        dqkv = torch.zeros(B, S, 3 * D, device="cuda", dtype=torch.bfloat16)
        dq, dk, dv = (dqkv[:, :, i * D:(i + 1) * D].view(B, S, H, d) for i in range(3))
        bias = None
    else:
        q = torch.randn(B, S, H, d, device="cuda").bfloat16()
        kv = torch.randn(B, L, 2 * D, device="cuda").bfloat16()
        k, v = kv[:, :, :D].view(B, L, H, d), kv[:, :, D:].view(B, L, H, d)
        dq = torch.zeros(B, S, H, d, device="cuda", dtype=torch.bfloat16)
        dkv = torch.zeros(B, L, 2 * D, device="cuda", dtype=torch.bfloat16)
        dk, dv = dkv[:, :, :D].view(B, L, H, d), dkv[:, :, D:].view(B, L, H, d)
        bias = torch.zeros(B, L, device="cuda")
    o = torch.empty(B, S, H, d, device="cuda", dtype=torch.bfloat16)
    return q, k, v, o, torch.empty(B, H, S, device="cuda"), torch.randn_like(o), dq, dk, dv, bias
for cross in (False, True):
    cs = [make(cross) for _ in range(NB)]
    for c in cs:
        ops.attn_fwd(c[0], c[1], c[2], c[3], d ** -0.5, bias=c[9], lse=c[4])
    for v1 in ("0", "1"):
        os.environ["TDM_B200_ATTN_V1"] = v1
        f = lambda i: ops.attn_fwd(cs[i % NB][0], cs[i % NB][1], cs[i % NB][2], cs[i % NB][3], d ** -0.5, bias=cs[i % NB][9], lse=cs[i % NB][4])
        a, _ = timeit(f, iters=50)
        g, mhz = timeit_graph(f, n=40)
        print(json.dumps({"kernel": "attn_fwd " + ("one-tile" if v1 == "1" else "two-stream"), "case": "cross" if cross else "self",
                          "us_python_loop": round(a, 2), "us_cuda_graph": round(g, 2), "sm_mhz": mhz}), flush=True)
    os.environ["TDM_B200_ATTN_V1"] = "0"
    f = lambda i: ops.attn_bwd(cs[i % NB][0], cs[i % NB][1], cs[i % NB][2], cs[i % NB][3], cs[i % NB][5], cs[i % NB][4], d ** -0.5,
                               cs[i % NB][6], cs[i % NB][7], cs[i % NB][8], bias=cs[i % NB][9])
    a, _ = timeit(f, iters=30)
    g, mhz = timeit_graph(f, n=30)
    print(json.dumps({"kernel": "attn_bwd (delta + main [+ cast])", "case": "cross" if cross else "self", "us_python_loop": round(a, 2),
                      "us_cuda_graph": round(g, 2), "sm_mhz": mhz}), flush=True)
