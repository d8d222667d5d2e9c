"""Fused global-norm clip + AdamW on a flat arena (tdm_sumsq_f32 / tdm_clip_coef / tdm_adamw_step, through
tdm_b200.ops) against torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW -- what the reference runs at
main.py:206-224 (AdamW betas 0.9/0.999, wd 1e-2, eps 1e-8, args.py:204-213) and main.py:536-540 (clip, step)."""
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def _ours_step(p, g, m, v, p16, t, lr, wd, max_norm, inv_scale, scratch):
    from tdm_b200 import ops
    sumsq, coef, norm = scratch
    ops.sumsq(g, sumsq, zero_first=True)
    ops.clip_coef(sumsq, max_norm, inv_scale, coef, norm)
    ops.adamw_step(p, g, m, v, p16, lr, 0.9, 0.999, 1e-8, wd, t, coef)
    return norm


@pytest.mark.parametrize("max_norm", [1.0, 0.0, 1e6])
@pytest.mark.parametrize("n", [4 * 1000, 4 * 262_147])
def test_adamw_and_clip_match_torch(n, max_norm):
    torch.manual_seed(0)
    lr, wd = 1e-3, 1e-2
    p0 = torch.randn(n, device="cuda")
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref_p], lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    p = p0.clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    p16 = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    scratch = (torch.zeros((), device="cuda", dtype=torch.float64), torch.ones((), device="cuda"),
               torch.zeros((), device="cuda"))
    for t in range(1, 6):
        g = torch.randn(n, device="cuda") * (0.05 if t % 2 else 3e-4)  # norms above and below the clip threshold
        ref_p.grad = g.clone()
        ref_norm = torch.nn.utils.clip_grad_norm_([ref_p], max_norm) if max_norm else g.norm()
        opt.step()
        norm = _ours_step(p, g, m, v, p16, t, lr, wd, max_norm, 1.0, scratch)
        assert abs(norm.item() - ref_norm.item()) <= 2e-6 * ref_norm.item()
        # fp32 op-order differences only (torch: addcdiv with a separately rounded denominator)
        err = (p - ref_p.detach()).abs().max().item()
        assert err <= 2e-6 * p0.abs().max().item() + 1e-4 * lr, (t, err)
        st = opt.state[ref_p]
        assert (m - st["exp_avg"]).abs().max().item() <= 1e-6 * st["exp_avg"].abs().max().item() + 1e-12
        assert (v - st["exp_avg_sq"]).abs().max().item() <= 1e-6 * st["exp_avg_sq"].abs().max().item() + 1e-16
        assert torch.equal(p16, p.bfloat16())  # the bf16 copy the tensor cores read is refreshed in the same pass


def test_world_scale_folds_the_average_into_the_clip_coefficient():
    """Data parallel: the arenas hold the SUM over ranks; clip + step with inv_scale = 1/world must equal clip + step on
    the averaged gradients (main.py:535-537 under DDP)."""
    torch.manual_seed(1)
    n, world = 4 * 5000, 4
    g = torch.randn(n, device="cuda") * 0.03
    outs = []
    for scale, inv in ((1.0, 1.0), (float(world), 1.0 / world)):
        p = torch.linspace(-1, 1, n, device="cuda")
        m, v = torch.zeros_like(p), torch.zeros_like(p)
        scratch = (torch.zeros((), device="cuda", dtype=torch.float64), torch.ones((), device="cuda"),
                   torch.zeros((), device="cuda"))
        norm = _ours_step(p, g * scale, m, v, None, 1, 1e-3, 1e-2, 1.0, inv, scratch)
        outs.append((p, norm.item()))
    assert abs(outs[0][1] - outs[1][1]) <= 1e-6 * outs[0][1]
    assert (outs[0][0] - outs[1][0]).abs().max().item() <= 1e-7


def test_flat_adamw_matches_torch_on_a_model():
    """FlatAdamW over a tiny DiT's arenas == torch.optim.AdamW over the same parameters, 3 steps with clipping."""
    from tdm_b200 import harness
    from tdm_b200.training import FlatAdamW
    student, fake, _ = harness.make_models("tiny", seed=0, device="cuda", lora_rank=0, teacher=False)
    opt = FlatAdamW(fake, lr=1e-3, weight_decay=1e-2, max_grad_norm=1.0)
    ref_p = torch.nn.Parameter(fake._flat.detach().clone())
    ref = torch.optim.AdamW([ref_p], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    gen = torch.Generator(device="cuda").manual_seed(3)
    for t in range(3):
        g = torch.randn(fake._grad.shape, device="cuda", generator=gen) * 1e-2
        fake._grad.copy_(g)
        ref_p.grad = g.clone()
        rn = torch.nn.utils.clip_grad_norm_([ref_p], 1.0)
        ref.step()
        opt.step()
        assert abs(opt.grad_norm.item() - rn.item()) <= 2e-6 * rn.item()
        assert (fake._flat - ref_p.detach()).abs().max().item() <= 2e-6 * ref_p.detach().abs().max().item() + 1e-7
        assert torch.equal(fake._flat16, fake._flat.bfloat16())
