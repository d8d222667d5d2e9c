"""Fixed cost of one tdm_gemm_bf16 launch: LoRA-sized problems against K (how much of their ~13 us is the launch floor?)."""
import json, sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "scripts")
from _timing import timeit, timeit_graph
from tdm_b200 import ops
NB = 3
for M, N, K in ((128, 64, 64), (8192, 64, 64), (8192, 64, 256), (8192, 64, 1152), (8192, 64, 4608), (8192, 192, 1152), (8192, 1152, 64),
                (960, 2304, 1152), (8192, 1152, 1152)):
    a = [torch.randn(M, K, device="cuda").bfloat16() for _ in range(NB)]
    b = [torch.randn(N, K, device="cuda").bfloat16() for _ in range(NB)]
    o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    us, mhz = timeit(lambda i: ops.gemm(a[i % NB], b[i % NB], out_bf16=o), iters=100)
    usg, _ = timeit_graph(lambda i: ops.gemm(a[i % NB], b[i % NB], out_bf16=o), n=60)
    print(json.dumps({"M": M, "N": N, "K": K, "us_python_loop": round(us, 2), "us_cuda_graph": round(usg, 2), "sm_mhz": mhz}), flush=True)
# empty-kernel floor of the stream (PDL launch of a trivial kernel): cast of 1 element
x = torch.zeros(8, device="cuda")
us, mhz = timeit(lambda i: ops.cast_bf16(x), iters=200)
print(json.dumps({"case": "cast_bf16 of 8 floats (launch floor incl. Python)", "us": round(us, 2)}), flush=True)
