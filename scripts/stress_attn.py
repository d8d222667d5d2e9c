"""Many attention-forward launches (self / cross, batch 8 / 16) interleaved with other work: looks for rare deadlocks."""
import sys, torch
sys.path.insert(0, ".")
from tdm_b200 import ops
def make(B, H, Sq, Sk, d):
    D = H * d
    if Sq == Sk:
        qkv = torch.randn(B, Sq, 3 * D, device="cuda").bfloat16()
        q, k, v = (qkv[:, :, i * D:(i + 1) * D].view(B, Sq, H, d) for i in range(3))
    else:
        q = torch.randn(B, Sq, H, d, device="cuda").bfloat16()
        kv = torch.randn(B, Sk, 2 * D, device="cuda").bfloat16()
        k, v = kv[:, :, :D].view(B, Sk, H, d), kv[:, :, D:].view(B, Sk, H, d)
    o = torch.empty(B, Sq, H, d, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, H, Sq, device="cuda")
    return q, k, v, o, lse
cases = [make(8, 16, 1024, 1024, 72), make(16, 16, 1024, 1024, 72), make(8, 16, 1024, 120, 72), make(16, 16, 1024, 120, 72)]
junk = torch.empty(1 << 26, device="cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
for it in range(n):
    for (q, k, v, o, lse) in cases:
        ops.attn_fwd(q, k, v, o, 72 ** -0.5, lse=lse)
    if it % 7 == 0:
        junk.zero_()  # flush L2 now and then: cold loads
    if it % 100 == 0:
        torch.cuda.synchronize(); print("it", it, flush=True)
torch.cuda.synchronize()
ref = torch.nn.functional.scaled_dot_product_attention(cases[0][0].transpose(1, 2).float(), cases[0][1].transpose(1, 2).float(), cases[0][2].transpose(1, 2).float()).transpose(1, 2)
print("ok, max err", (cases[0][3].float() - ref).abs().max().item())
