"""As the tests: fresh model, CPU oracle forward in between (GPU idles), then the GPU forward twice; diff every saved buffer."""
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from test_dit_gpu import _pair, _inputs, TINY
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
lora = len(sys.argv) > 2 and sys.argv[2] == "lora"
names = "mod mean1 rstd1 mean2 rstd2 lse1 lse2 x1 x2 h1 qkv a1 y1 x1b q2 kv2 a2 h2 u g y3 t_qkv t_o t_q2 t_kv2 t_o2 t_ff1 t_ff2".split()
def sizes(B, N, D, H, L, Dff, r):
    M, BL = B * N, B * L
    al = lambda b: (b + 255) & ~255
    f = [B * 6 * D, M, M, M, M, B * H * N, B * H * N, M * D, M * D]
    h = [M * D, M * 3 * D, M * D, M * D, M * D, M * D, BL * 2 * D, M * D, M * D, M * Dff, M * Dff, M * D,
         M * 3 * r, M * r, M * r, BL * 2 * r, M * r, M * r, M * r]
    return [al(x * 4) for x in f] + [al(x * 2) for x in h]
bad_total = 0
for it in range(n):
    ref, ours = _pair(TINY, lora=lora)
    lat, t, ctx, mask = _inputs(TINY, 2, 120, seed=it)
    with torch.no_grad():
        out_ref = ref(lat, timestep=t, encoder_hidden_states=ctx, encoder_attention_mask=mask)[0]
    args = (lat.cuda(), t.cuda(), ctx.cuda(), mask.cuda())
    with torch.no_grad():
        o1, s1 = ours._forward_impl(*args, True)
        torch.cuda.synchronize()
        o2, s2 = ours._forward_impl(*args, True)
    torch.cuda.synchronize()
    if not torch.equal(o1, o2):
        bad_total += 1
        print("iter", it, "first forward differs from second: rel", ((o1 - o2).abs().max() / o2.abs().max()).item())
        for nm in ("patches", "tsin", "t_h", "semb", "tmod", "c1", "ctx"):
            a, b = getattr(s1, nm), getattr(s2, nm)
            if a is not None and not torch.equal(a, b):
                print("   pre-block buffer differs:", nm, ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-9)).item())
        sh = s1.shape
        sz = sizes(sh.B, sh.N, sh.D, sh.H, sh.L, sh.Dff, sh.lora_r if sh.lora_r else 0)
        for bi, (w1, w2) in enumerate(zip(s1.ws, s2.ws)):
            if not torch.equal(s1.xs[bi], s2.xs[bi]):
                print("   block", bi, "INPUT x differs")
            off = 0
            for nm, z in zip(names, sz):
                if z and not torch.equal(w1[off:off + z], w2[off:off + z]):
                    print("   block", bi, "buffer", nm, "differs (bytes", z, ")")
                    if nm == "a1" and bi == 0:
                        M, D = sh.B * sh.N, sh.D
                        A1 = w1[off:off + M * D * 2].view(torch.bfloat16).view(M, D).float()
                        A2 = w2[off:off + M * D * 2].view(torch.bfloat16).view(M, D).float()
                        dd = (A1 - A2).abs()
                        rows = (dd.max(dim=1).values > 0).nonzero().flatten().tolist()
                        cols = (dd.max(dim=0).values > 0).nonzero().flatten().tolist()
                        print("      a1 rows differing:", len(rows), rows[:6], "..", rows[-3:], " cols differing:", len(cols), cols[:6], "..", cols[-3:])
                        print("      max |a1_first| %.3g max |a1_second| %.3g maxdiff %.3g; H=%d" % (A1.abs().max(), A2.abs().max(), dd.max(), sh.H))
                        r0 = rows[0]
                        print("      row", r0, "first:", A1[r0, cols[0]:cols[0] + 6].tolist(), "second:", A2[r0, cols[0]:cols[0] + 6].tolist())
                off += z
            break_after = True
            if bi >= 1: break
    del ours, ref
print("done, bad iterations:", bad_total)
