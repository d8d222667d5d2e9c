"""Run-to-run determinism of the tiny LoRA forward / backward with NaN-poisoned allocator blocks (debug aid)."""
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from test_dit_gpu import _pair, _inputs, TINY
ref, ours = _pair(TINY, lora=True)
lat, t, ctx, mask = _inputs(TINY, 2, 24)
args = (lat.cuda(), t.cuda(), ctx.cuda(), mask.cuda())
def poison():
    xs = [torch.full((n,), float("nan"), device="cuda") for n in (1 << 24, 1 << 22, 1 << 20, 1 << 18, 1 << 16, 3 << 14, 5 << 12, 1 << 10)] * 2
    del xs
base = None
gout = torch.randn(2, 4, TINY.sample_size, TINY.sample_size, generator=torch.Generator().manual_seed(5)).cuda()
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
    poison()
    with torch.no_grad():
        om, _ = ours._forward_impl(*args, False)
        of, _ = ours._forward_impl(*args, True)
    for p in ours.parameters():
        p.grad = None
    ours._grad = None if hasattr(ours, "_grad") and False else getattr(ours, "_grad", None)
    out = ours(args[0], timestep=args[1], encoder_hidden_states=args[2], encoder_attention_mask=args[3])[0]
    if ours._lgrad is not None: ours._lgrad.zero_()
    (out.chunk(2, dim=1)[0] * gout).sum().backward()
    g = ours._lgrad.clone()
    torch.cuda.synchronize()
    cur = (om.clone(), of.clone(), g)
    if base is None:
        base = cur
        print("base finite:", [bool(torch.isfinite(x).all()) for x in base])
    else:
        bad = [i for i, (a, b) in enumerate(zip(base, cur)) if not torch.equal(a, b)]
        if bad:
            print("iter", it, "mismatch in", bad, [((a - b).abs().max().item(), bool(torch.isfinite(b).all())) for a, b in zip(base, cur)])
print("done")
