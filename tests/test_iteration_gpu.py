"""The repaired TDM iteration on the GPU (tdm_b200.training.TdmTrainer) against the fp32 CPU oracle
(oracle/iteration.RepairedIteration): same weights, same seeds, same synthetic T5 embeddings.

* timestep / index selection: bit-exact (north_star);
* every intermediate of one iteration within bf16 tolerance of the fp32 path;
* the TDM loss within 1 % on EVERY one of 100 iterations of BASELINE.json configs[0] (tiny DiT: 2 blocks, hidden 256,
  32x32x4 latent, batch 2), optimiser steps applied on both sides.

Measured on B200 (scripts/parity_stats.py -> profiles/r02_parity_stats.json): TDM loss deviation max 0.18 %, median
0.003 %; intermediates 0.5-0.8 % of the largest entry (0.14-0.17 % of the mean magnitude); sign of the Huber gradient
equal on 99.2 % of the entries."""
import pytest
import torch

from oracle.dit import OracleDiT, TINY
from oracle import lora as oracle_lora
from oracle.iteration import RepairedIteration
from oracle.schedule import OracleDDPMScheduler

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1800)]

LOWP = torch.bfloat16


def _rel(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return ((got - ref).abs().max() / (ref.abs().max() + 1e-12)).item()


def _mean_rel(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return ((got - ref).abs().mean() / (ref.abs().mean() + 1e-12)).item()


def _build(seed=0, lora=True, lr=2e-5):
    from tdm_b200 import harness
    from tdm_b200.training import TdmTrainer
    from tdm_b200.transformer import Transformer2DModel
    torch.manual_seed(seed)
    base = OracleDiT(TINY)
    with torch.no_grad():  # non-trivial modulation / gates (default init leaves them tiny)
        base.adaln_single.linear.weight.mul_(3.0)
        for blk in base.transformer_blocks:
            blk.scale_shift_table.mul_(8.0)
    sd = base.state_dict()
    o_student, o_fake, o_teacher = OracleDiT(TINY), OracleDiT(TINY), OracleDiT(TINY)
    for m in (o_student, o_fake, o_teacher):
        m.load_state_dict(sd)
    # In the reference all three start from one checkpoint, so x0_fake - x0_teacher is at first only the
    # cond/uncond difference of ONE network: a small difference of two nearly equal bf16 results whose sign
    # (the Huber gradient) is noise.  Detune the fake model so the score difference is well conditioned.
    with torch.no_grad():
        o_fake.proj_out.weight.mul_(1.3)
        o_fake.proj_out.bias.add_(0.05)
    if lora:
        oracle_lora.add_lora(o_student, r=64, alpha=64)
        with torch.no_grad():
            g = torch.Generator().manual_seed(seed + 7)
            for n, p in o_student.named_parameters():
                if ".lora_B." in n:
                    p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    cfg = TINY.to_dict()
    student = Transformer2DModel(cfg, device="cuda", seed=1)
    fake = Transformer2DModel(cfg, device="cuda", init_weights=False)
    teacher = Transformer2DModel(cfg, device="cuda", init_weights=False)
    if lora:
        student.add_lora(rank=64, alpha=64)
    student.load_oracle_state(o_student.state_dict())
    fake.load_oracle_state(o_fake.state_dict())
    teacher.load_oracle_state(sd)
    teacher.requires_grad_(False)
    B, L = 2, 24
    ctx, mask, uctx, umask = harness.synthetic_text(B, L=L, Cc=TINY.caption_channels, seed=seed)
    kw = dict(cfg=4.5, total_steps=900, use_huber=True, use_separate=True, lowp=LOWP, lr=lr, lr_fake=lr)
    oracle = RepairedIteration(o_student, o_fake, o_teacher, OracleDDPMScheduler(), uctx, umask, **kw)
    ours = TdmTrainer(student, fake, teacher, harness.make_scheduler(), uctx.cuda(), umask.cuda(), **kw)
    return oracle, ours, (ctx, mask)


def _draws(seed, B, S, total_steps=900, sep=True):
    """The reference's draw sequence on the CPU generator (device generators differ between CPU and CUDA by
    construction, so both sides are fed the same host-drawn values)."""
    from tdm_b200.training import draw_fake_branch
    torch.manual_seed(seed)
    noise = torch.randn([B, 4, S, S]).to(LOWP)
    ind_t, t_fake, noise_fake = draw_fake_branch(noise, total_steps, sep)
    return noise, (ind_t, t_fake, noise_fake)


def test_one_iteration_intermediates_and_gradients():
    oracle, ours, (ctx, mask) = _build()
    noise, draws = _draws(11, 2, TINY.sample_size)
    ro = oracle.step(noise, ctx, mask, draws=draws, apply=False)
    # product side: same values, on the device; stop before the optimizer by reading gradients after step
    # (lr = 0 keeps weights while still running the fused clip + AdamW kernels)
    for opt in (ours.opt_student, ours.opt_fake):
        opt.lr = 0.0
        opt.wd = 0.0
    kept = {}
    zs, zf = ours.opt_student.zero_grad, ours.opt_fake.zero_grad
    ours.opt_student.zero_grad = lambda: kept.__setitem__("s", [g.clone() for _, g, _ in ours.opt_student.groups]) or zs()
    ours.opt_fake.zero_grad = lambda: kept.__setitem__("f", [g.clone() for _, g, _ in ours.opt_fake.groups]) or zf()
    rg = ours.step(noise.cuda(), ctx.cuda(), mask.cuda(), draws=tuple(d.cuda() for d in draws))
    torch.cuda.synchronize()
    # integer selections: bit-exact
    for key in ("ind_t", "t_fake", "t1", "t_in"):
        assert torch.equal(ours.last[key].cpu(), oracle.last[key]), key
    assert abs(rg["loss"].item() - ro["loss"].item()) <= 2e-3 * abs(ro["loss"].item())
    assert abs(rg["loss_fake"].item() - ro["loss_fake"].item()) <= 1e-2 * abs(ro["loss_fake"].item())
    # every intermediate of the iteration: bf16 tensor-core DiT against the fp32 oracle.
    # max|a-b| / max|b| <= 2e-2 (north_star's norm) AND mean|a-b| / mean|b| <= 5e-3 (no hiding behind one large entry)
    for key in ("x_in", "x0_sel", "latents_fake", "x0_fake", "sd", "sdu"):
        a, b = ours.last[key], oracle.last[key]
        assert a.shape == b.shape, key
        assert _rel(a, b) < 2e-2, (key, _rel(a, b))
        assert _mean_rel(a, b) < 5e-3, (key, _mean_rel(a, b))
    # eps_target = (x_t - alpha x0_sel) / sigma at t2 subtracts nearly equal quantities and divides by sigma_t2 (0.01 at
    # t2 ~ 0): a 0.5 % difference of x0_sel is amplified by alpha/sigma (up to 67 for these draws), so only its mean
    # error is meaningful
    assert _mean_rel(ours.last["eps_target"], oracle.last["eps_target"]) < 5e-2
    # closed-form Huber gradient (c = 1e-3): ~ sign(x0_fake - target) / (w_b n); the sign is noise where the difference
    # is below bf16 resolution
    g, gr = ours.last["g_x0"].cpu(), oracle.last["g_x0"].float()
    assert (torch.sign(g) == torch.sign(gr)).float().mean().item() > 0.98
    assert _mean_rel(g, gr) < 4e-2
    # student LoRA gradients and fake-score full gradients against autograd
    s_named = dict(ours.student.named_parameters())
    worst = []
    lgrad = kept["s"][-1]
    for n, p in oracle.student.named_parameters():
        if p.grad is None or p.grad.abs().max() == 0:
            continue
        got = ours.student._lora_arena.view(lgrad, n)
        worst.append((_rel(got, p.grad), "student." + n))
    fgrad = kept["f"][0]
    for n, p in oracle.fake.named_parameters():
        if p.grad is None or p.grad.abs().max() == 0 or n.endswith("to_k.bias"):
            continue
        got = ours.fake._arena.view(fgrad, n)
        worst.append((_rel(got, p.grad), "fake." + n))
    worst.sort(reverse=True)
    assert len(worst) > 40
    # bf16 gradient operands against fp32 autograd; measured on B200 (profiles/r02_parity_stats.json): worst 3.8e-2,
    # mean 1.3e-2, median 0.9e-2 of the largest entry of each gradient tensor
    assert worst[0][0] < 5e-2, worst[:6]
    assert sum(w for w, _ in worst) / len(worst) < 2e-2
    # the fused clip kernel saw the same global norm torch computes
    ref_norm = torch.nn.utils.clip_grad_norm_([p for p in oracle.fake.parameters() if p.grad is not None], 1e9)
    assert abs(ours.opt_fake.grad_norm.item() - ref_norm.item()) <= 3e-2 * ref_norm.item()


def test_losses_track_the_oracle_over_many_iterations():
    """BASELINE.json configs[0] / north_star: loss within 1 % over 100 iterations -- asserted on EVERY iteration for the
    TDM loss (main.py:514-529), identical seeds and draws on both sides, optimiser steps applied on both sides.

    The auxiliary fake-score MSE (repair D5; not a quantity of the reference) cannot meet 1 % per iteration in bf16 and
    the test says so with numbers: loss_fake = mean((eps_fake - eps_target)^2) is the square of a RESIDUAL of two
    O(1) quantities, so a bf16 forward error of ~0.5 % of eps becomes 2 * 0.005 * |eps| / |residual| ~ 1-2 % of the
    loss, and eps_target itself divides by sigma_t2 (alpha/sigma up to 67 in these draws).  Measured on B200: median
    0.37 %, 21 of 100 iterations above 1 %, max 4.1 % (iteration 48, t2 = 10); the mean over the run agrees to 0.4-1.3 %
    (it is dominated by a few iterations with losses of 20-50).  The worst iteration moves from run to run (4.1-6.3 % seen):
    split-K partial sums and the dQ tiles are ADDED in HBM in arrival order (TMA reduce-add), so two runs differ in the
    last fp32 bit of some gradients and, 48 optimiser steps later, in that one ill-conditioned loss value; its bound is
    10 %, the median bound stays 1 %."""
    n_iter = 100
    oracle, ours, (ctx, mask) = _build(seed=3, lr=1e-4)
    lo, lg, fo, fg = [], [], [], []
    for it in range(n_iter):
        noise, draws = _draws(1000 + it, 2, TINY.sample_size)
        ro = oracle.step(noise, ctx, mask, draws=draws)
        rg = ours.step(noise.cuda(), ctx.cuda(), mask.cuda(), draws=tuple(d.cuda() for d in draws))
        lo.append(ro["loss"].item()); lg.append(rg["loss"].item())
        fo.append(ro["loss_fake"].item()); fg.append(rg["loss_fake"].item())
    lo, lg, fo, fg = map(torch.tensor, (lo, lg, fo, fg))
    assert torch.isfinite(lg).all() and torch.isfinite(fg).all()
    assert ((lg - lo).abs() / lo.abs()).max() <= 1e-2          # north_star: every iteration within 1 %
    assert abs(lg.mean() - lo.mean()) <= 2e-3 * lo.mean().abs()
    dev_f = (fg - fo).abs() / fo.abs()
    # heavy tail (a handful of iterations with t2 ~ 0 carry losses of 20-50 and dominate the mean): 2 % on the mean, 1 % on
    # the typical iteration
    assert abs(fg.mean() - fo.mean()) <= 2e-2 * fo.mean().abs()
    assert dev_f.median() < 1e-2 and dev_f.max() < 1e-1


def test_sampling_4nfe_matches_oracle():
    from tdm_b200 import harness
    from tdm_b200.training import sample_4nfe, SAMPLING_MID_POINTS
    from oracle import glue
    oracle, ours, (ctx, mask) = _build(seed=5)
    noise, _ = _draws(21, 2, TINY.sample_size)
    sch = OracleDDPMScheduler()
    with torch.no_grad():
        ref = glue.few_step_sample(oracle.student, sch, noise, noise, ctx, mask, steps=4, mid_points=SAMPLING_MID_POINTS,
                                   total_steps=1000, lowp=LOWP)
    got = sample_4nfe(ours.student, harness.make_scheduler(), noise.cuda(), ctx.cuda(), mask.cuda(), lowp=LOWP)
    assert got.dtype == torch.float32 and got.shape == ref.shape
    assert _rel(got, ref) < 3e-2


def test_deferred_fake_optimizer_gives_the_same_iterations():
    """TdmTrainer(defer_fake_opt=True) runs the fake-score optimiser step of iteration i on the side stream under the
    trajectory of iteration i+1: losses and both models' weights after sync() must match the in-order run (same kernels,
    same order per tensor; fp32 atomics in the backward reductions make two runs differ in the last bits only)."""
    from tdm_b200 import harness
    from tdm_b200.training import TdmTrainer
    results = []
    for defer in (False, True):
        torch.manual_seed(7)
        torch.cuda.manual_seed(7)
        student, fake, teacher = harness.make_models("tiny", seed=0, device="cuda", lora_rank=64)
        ctx, mask, uctx, umask = harness.synthetic_text(2, seed=5, Cc=student.Cc)
        tr = TdmTrainer(student, fake, teacher, harness.make_scheduler(), uctx.cuda(), umask.cuda(), lr=1e-3, lr_fake=1e-3,
                        defer_fake_opt=defer)
        S = student.sample_size
        losses = []
        for i in range(4):
            torch.manual_seed(100 + i)
            noise = torch.randn(2, 4, S, S).to(torch.bfloat16).cuda()
            out = tr.step(noise, ctx.cuda(), mask.cuda())
            losses.append((out["loss"].item(), out["loss_fake"].item()))
        tr.sync()
        torch.cuda.synchronize()
        results.append((losses, fake._flat.clone(), student._lflat.clone()))
    for (l0, f0), (l1, f1) in zip(results[0][0], results[1][0]):
        assert abs(l0 - l1) <= 1e-3 * abs(l0) + 1e-6 and abs(f0 - f1) <= 1e-3 * abs(f0) + 1e-6
    assert _rel(results[1][1], results[0][1]) < 1e-4   # fake-score weights after four optimiser steps
    assert _rel(results[1][2], results[0][2]) < 1e-3   # student LoRA weights
