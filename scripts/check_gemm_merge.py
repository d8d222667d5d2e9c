import sys, torch
sys.path.insert(0, ".")
from tdm_b200 import ops
def run(M, N, K, b_mn=True, feat="merge"):
    torch.manual_seed(3)
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = torch.randn((K, N) if b_mn else (N, K), device="cuda").bfloat16()
    res = torch.randn(M, N, device="cuda")
    ref = 2.0 * (a.float() @ (b.float() if b_mn else b.float().t())) + res
    for pr, tn in ((1, 0), (0, 0), (2, 256), (2, 128)):
        if feat == "merge":
            o = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
            ops.gemm(a, b, b_mn_major=b_mn, alpha=2.0, res=res, out_bf16=o, cta_pair=pr, tile_n=tn)
        else:
            o = torch.zeros(M, N, device="cuda")
            ops.gemm(a, b, b_mn_major=b_mn, alpha=2.0, res=res, out_f32=o, cta_pair=pr, tile_n=tn)
        d = (o.float() - ref).abs()
        print(f"{feat} M={M} N={N} K={K} pair={pr} tn={tn}: maxdiff {d.max().item():.4g} at {divmod(d.argmax().item(), N)} (ref max {ref.abs().max().item():.3g})")
for feat in ("merge", "res32"):
    run(256, 256, 64, feat=feat); run(1152, 1152, 64, feat=feat); run(1152, 4608, 64, feat=feat); run(512, 512, 256, b_mn=False, feat=feat)
