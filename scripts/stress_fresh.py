"""Fresh model per iteration (as the tests do): forward / backward results must be bit-identical run to run."""
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from test_dit_gpu import _pair, _inputs, TINY
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
lora = len(sys.argv) > 2 and sys.argv[2] == "lora"
base = None
bad_total = 0
for it in range(n):
    ref, ours = _pair(TINY, lora=lora)
    lat, t, ctx, mask = _inputs(TINY, 2, 120)
    args = (lat.cuda(), t.cuda(), ctx.cuda(), mask.cuda())
    with torch.no_grad():
        o1, saved = ours._forward_impl(*args, True)
        o2, _ = ours._forward_impl(*args, False)
    xs = [x.clone() for x in saved.xs]
    out = ours(args[0], timestep=args[1], encoder_hidden_states=args[2], encoder_attention_mask=args[3])[0]
    out.sum().backward()
    g = (ours._lgrad if lora else ours._grad).clone()
    torch.cuda.synchronize()
    cur = [o1.clone(), o2.clone(), g] + xs
    if base is None:
        base = cur
    else:
        bad = [i for i, (a, b) in enumerate(zip(base, cur)) if not torch.equal(a, b)]
        if bad:
            bad_total += 1
            print("iter", it, "mismatch idx", bad, ["%.3g" % ((a.float() - b.float()).abs().max().item() / (a.float().abs().max().item() + 1e-9)) for a, b in zip(base, cur)])
    del ours, ref
print("done, bad iterations:", bad_total)
