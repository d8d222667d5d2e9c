"""auto tile choice vs forced single-CTA on odd shapes (debug aid)."""
import sys, torch
sys.path.insert(0, ".")
from tdm_b200 import ops
torch.manual_seed(0)
def run(M, N, K, a_mn, b_mn, acc=False, seg=None, **kw):
    a = torch.randn((K, M) if a_mn else (M, K), device="cuda").bfloat16()
    b = torch.randn((K, N) if b_mn else (N, K), device="cuda").bfloat16()
    outs = []
    for pr in (1, 0, 2):
        extra = {}
        if seg:
            nseg, r = seg
            torch.manual_seed(1)
            t = torch.randn(M, nseg * r, device="cuda").bfloat16(); bl = torch.randn(N, r, device="cuda").bfloat16()
            extra = dict(a2=t, b2=bl, seg_n=N // nseg, seg_k2=r)
        try:
            if acc:
                o = torch.ones(M, N, device="cuda")
                ops.gemm(a, b, a_mn_major=a_mn, b_mn_major=b_mn, res=o, out_f32=o, cta_pair=pr, **extra, **kw)
            else:
                o = torch.empty(M, N, device="cuda")
                ops.gemm(a, b, a_mn_major=a_mn, b_mn_major=b_mn, out_f32=o, cta_pair=pr, **extra, **kw)
            outs.append(o)
        except Exception as e:
            outs.append(None); print("   err", pr, str(e)[-80:])
    ref = outs[0]
    for name, o in zip(("auto", "pair"), outs[1:]):
        if o is not None:
            print(f"M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} acc={acc} seg={seg} {name}: maxdiff {(o - ref).abs().max().item():.4g} (ref max {ref.abs().max().item():.3g})")
run(192, 256, 512, 1, 1, acc=True)
run(256, 64, 512, 1, 1, acc=True)
run(768, 256, 512, 1, 1, acc=True)
run(512, 768, 256, 0, 0, seg=(3, 64))
run(512, 256, 768, 0, 1)
run(512, 256, 768 , 0, 1, seg=(1, 64))
run(240, 512, 256, 0, 0, seg=(2, 64))
run(512, 1024, 256, 0, 1)
print("--- K2 pairs")
def run2(M, N, K, K2, b_mn, **kw):
    torch.manual_seed(2)
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = torch.randn((K, N) if b_mn else (N, K), device="cuda").bfloat16()
    a2 = torch.randn(M, K2, device="cuda").bfloat16()
    b2 = torch.randn((K2, N) if b_mn else (N, K2), device="cuda").bfloat16()
    outs = []
    for pr in (1, 0, 2):
        o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        ops.gemm(a, b, b_mn_major=b_mn, a2=a2, b2=b2, out_bf16=o, cta_pair=pr, **kw)
        outs.append(o.float())
    ref = (a.float() @ (b.float() if b_mn else b.float().t()) + a2.float() @ (b2.float() if b_mn else b2.float().t()))
    for name, o in zip(("single", "auto", "pair"), outs):
        print(f"M={M} N={N} K={K} K2={K2} b_mn={b_mn} {name}: maxdiff vs fp32 {(o - ref).abs().max().item():.4g} (ref max {ref.abs().max().item():.3g})")
run2(512, 256, 768, 192, 1)
run2(512, 256, 256, 64, 1)
run2(512, 1024, 256, 64, 1)
run2(512, 256, 1024, 64, 1)
run2(512, 256, 256, 64, 0)
run2(512, 1024, 256, 64, 0)
run2(240, 256, 256, 64, 1)
