"""Accumulate-only wgrad GEMMs (dW += dy^T x, tokens = 8192) against the split-K factor: calibrates gemm_plan_splits now that
the partial sums are added by TMA reduce operations instead of per-thread vector reds."""
import json, sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "scripts")
from _timing import timeit
from tdm_b200 import ops
NB = 3
for name, tokens, n_out, k_in in (("w_o", 8192, 1152, 1152), ("w_qkv", 8192, 3456, 1152), ("w_ff1", 8192, 4608, 1152),
                                  ("w_ff2", 8192, 1152, 4608), ("w_kv2", 960, 2304, 1152), ("lora_dB", 8192, 1152, 64),
                                  ("lora_dA", 8192, 64, 1152), ("lora_dA3", 8192, 192, 1152), ("w_o_4096tok", 32768, 1152, 1152),
                                  ("w_ff1_4096tok", 32768, 4608, 1152)):
    dy = [torch.randn(tokens, n_out, device="cuda").bfloat16() for _ in range(NB)]
    x = [torch.randn(tokens, k_in, device="cuda").bfloat16() for _ in range(NB)]
    dw = torch.zeros(n_out, k_in, device="cuda")
    row = {"case": name, "tokens": tokens, "n_out": n_out, "k_in": k_in}
    for sp in (0, 1, 2, 3, 4, 5, 6, 8, 12):
        try:
            us, _ = timeit(lambda i: ops.gemm(dy[i % NB], x[i % NB], a_mn_major=True, b_mn_major=True, res=dw, out_f32=dw, split_k=sp), iters=30)
            row[f"split{sp}_us"] = round(us, 2)
        except Exception as e:  # noqa: BLE001
            row[f"split{sp}_us"] = str(e)[:40]
    print(json.dumps(row), flush=True)
