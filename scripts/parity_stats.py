"""Parity numbers behind the tolerances asserted in tests/ (run on the GPU box; writes gpurun_out/parity_stats.json):
  * one repaired iteration: every intermediate tensor against oracle/iteration.py (global max-relative and mean-relative);
  * 100 iterations of BASELINE configs[0]: per-iteration relative loss deviation (max / quantiles / offenders);
  * per-block outputs: global norm max|a-b|/max|b| next to a per-token-row norm and an elementwise norm with a floor.
"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def norms(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    d = (got - ref).abs()
    out = {"global": (d.max() / (ref.abs().max() + 1e-30)).item(),
           "mean_rel": (d.mean() / (ref.abs().mean() + 1e-30)).item()}
    if ref.dim() >= 2:
        r2, d2 = ref.reshape(-1, ref.shape[-1]), d.reshape(-1, ref.shape[-1])
        out["per_row"] = (d2.max(dim=1).values / (r2.abs().max(dim=1).values + 1e-30)).max().item()
    floor = 0.05 * ref.abs().max()
    out["elementwise_floor5pct"] = (d / (ref.abs() + floor)).max().item()
    return out


def main():
    import test_iteration_gpu as T
    from oracle.dit import TINY, DiTConfig
    res = {}
    oracle, ours, (ctx, mask) = T._build()
    noise, draws = T._draws(11, 2, TINY.sample_size)
    for opt in (ours.opt_student, ours.opt_fake):
        opt.lr = 0.0
        opt.wd = 0.0
    ro = oracle.step(noise, ctx, mask, draws=draws, apply=False)
    kept = {}
    zs, zf = ours.opt_student.zero_grad, ours.opt_fake.zero_grad
    ours.opt_student.zero_grad = lambda: kept.__setitem__("s", [g.clone() for _, g, _ in ours.opt_student.groups]) or zs()
    ours.opt_fake.zero_grad = lambda: kept.__setitem__("f", [g.clone() for _, g, _ in ours.opt_fake.groups]) or zf()
    rg = ours.step(noise.cuda(), ctx.cuda(), mask.cuda(), draws=tuple(d.cuda() for d in draws))
    ours.sync()
    torch.cuda.synchronize()
    gerr = []
    for n, p in oracle.student.named_parameters():
        if p.grad is not None and p.grad.abs().max() > 0:
            gerr.append((norms(ours.student._lora_arena.view(kept["s"][-1], n), p.grad)["global"], "student." + n))
    for n, p in oracle.fake.named_parameters():
        if p.grad is not None and p.grad.abs().max() > 0 and not n.endswith("to_k.bias"):
            gerr.append((norms(ours.fake._arena.view(kept["f"][0], n), p.grad)["global"], "fake." + n))
    gerr.sort(reverse=True)
    ge = torch.tensor([g for g, _ in gerr])
    res["iteration_gradients"] = {"n": len(gerr), "worst": gerr[:8], "mean": ge.mean().item(), "median": ge.median().item(),
                                  "p90": ge.quantile(0.9).item()}
    res["iteration_intermediates"] = {k: norms(ours.last[k], oracle.last[k]) for k in
                                      ("latents_fake", "x_in", "x0_sel", "x0_fake", "eps_target", "sd", "sdu", "g_x0")}
    res["iteration_losses"] = {k: [rg[k].item(), ro[k].item()] for k in ("loss", "loss_fake")}
    g, r = ours.last["g_x0"].cpu(), oracle.last["g_x0"]
    res["g_x0_sign_agreement"] = (torch.sign(g) == torch.sign(r)).float().mean().item()

    n_iter = 100
    oracle, ours, (ctx, mask) = T._build(seed=3, lr=1e-4)
    dev = {"loss": [], "loss_fake": []}
    vals = {"loss": [], "loss_fake": []}
    for it in range(n_iter):
        noise, draws = T._draws(1000 + it, 2, TINY.sample_size)
        ro = oracle.step(noise, ctx, mask, draws=draws)
        rg = ours.step(noise.cuda(), ctx.cuda(), mask.cuda(), draws=tuple(d.cuda() for d in draws))
        for k in dev:
            a, b = rg[k].item(), ro[k].item()
            dev[k].append(abs(a - b) / abs(b))
            vals[k].append((a, b))
        if it % 10 == 0:
            print(it, dev["loss"][-1], dev["loss_fake"][-1], flush=True)
    for k in dev:
        d = torch.tensor(dev[k])
        offenders = [(i, dev[k][i], vals[k][i], oracle_t) for i, oracle_t in enumerate([None] * n_iter) if dev[k][i] > 1e-2]
        res["loss100/" + k] = {"max": d.max().item(), "median": d.median().item(), "p90": d.quantile(0.9).item(),
                               "p99": d.quantile(0.99).item(), "n_over_1pct": int((d > 1e-2).sum()),
                               "mean_rel_of_means": abs(sum(a for a, _ in vals[k]) - sum(b for _, b in vals[k])) / abs(sum(b for _, b in vals[k])),
                               "offenders": [(i, dv, v) for i, dv, v, _ in offenders][:20]}

    import test_dit_gpu as D
    for name, cfg, B, L in (("tiny", TINY, 2, 120), ("xl_width_2blocks", DiTConfig(num_layers=2), 2, 120)):
        ref, mine = D._pair(cfg)
        lat, t, c, m = D._inputs(cfg, B, L)
        ref.capture = {}
        with torch.no_grad():
            out_ref = ref(lat, timestep=t, encoder_hidden_states=c, encoder_attention_mask=m)[0]
            out, saved = mine._forward_impl(lat.cuda(), t.cuda(), c.cuda(), m.cuda(), True)
        res["blocks/" + name] = {f"block{i}": norms(saved.xs[i + 1].view(B, -1, cfg.inner_dim), ref.capture[f"block{i}"])
                                 for i in range(cfg.num_layers)}
        res["blocks/" + name]["output"] = norms(out, out_ref)
    out = ROOT / "gpurun_out" / "parity_stats.json"
    out.parent.mkdir(exist_ok=True)
    out.write_text(json.dumps(res, indent=1))
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
