"""LN-modulate forward (M = 8192 rows of 1152 fp32 -> bf16, 56.6 MB): one-row-per-warp grid against the streaming form,
inputs rotating over 6 buffers (226 MB > L2) and, second, a single L2-resident input as in the iteration."""
import json, os, sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "scripts")
from _timing import timeit_graph
from tdm_b200 import ops
M, D, N = 8192, 1152, 1024
xs = [torch.randn(M, D, device="cuda") for _ in range(6)]
mod = torch.randn(8, 6, D, device="cuda")
for flag in ("0", "1", "0", "1"):
    os.environ["TDM_B200_LN_STREAM"] = flag
    us_cold, _ = timeit_graph(lambda i: ops.ln_modulate_fwd(xs[i % 6], mod[:, 0], mod[:, 1], N, 1e-6), n=60)
    us_hot, _ = timeit_graph(lambda i: ops.ln_modulate_fwd(xs[0], mod[:, 0], mod[:, 1], N, 1e-6), n=60)
    print(json.dumps({"stream": flag, "us_rotating_inputs": round(us_cold, 2), "us_same_input": round(us_hot, 2),
                      "GBps_rotating": round(56.6e6 / us_cold / 1e3, 1)}), flush=True)
