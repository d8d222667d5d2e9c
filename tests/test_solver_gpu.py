"""Fused solver / re-noise / loss kernels (through the C ABI) against the CPU oracle restatement of
the reference glue.  Elementwise results must be BIT-EXACT (same IEEE op sequence); the loss scalar is
an fp64 reduction whose summation order differs, tolerance 1e-12 relative (stated below)."""
import pytest
import torch

from oracle import glue
from oracle.schedule import OracleDDPMScheduler

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def _tables(dtype):
    sch = OracleDDPMScheduler()
    a = torch.sqrt(sch.alphas_cumprod).to(dtype)
    s = torch.sqrt(1 - sch.alphas_cumprod).to(dtype)
    return sch, a, s


@pytest.mark.parametrize("lowp", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("with_cfg", [False, True])
@pytest.mark.parametrize("pred", ["epsilon", "v_prediction"])
def test_solver_step_bit_exact(lowp, with_cfg, pred):
    from tdm_b200 import ops
    torch.manual_seed(0)
    B, C, H, W = 3, 4, 16, 16
    sch, a, s = _tables(lowp)
    out_c = torch.randn(B, 2 * C, H, W)
    out_u = torch.randn(B, 2 * C, H, W) if with_cfg else None
    x = torch.randn(B, C, H, W).to(lowp)
    t = torch.tensor([899, 674, 3])
    tn = torch.tensor([674, 449, -1])  # -1 wraps to the last table entry, like models.py:57-58 after the last step
    cfg = 4.5
    # oracle (CPU)
    eps = out_c.chunk(2, dim=1)[0]
    if with_cfg:
        eu = out_u.chunk(2, dim=1)[0]
        eps = eu + cfg * (eps - eu)
    x0_ref = glue.x0_from_model_output(eps, t, x, pred, a, s)
    xn_ref = sch.add_noise(x0_ref, eps, tn).to(lowp)
    d = "cuda"
    e, x0, xn = ops.solver_step(out_c.to(d), x.to(d), t.to(d), a.to(d), s.to(d), pred,
                                model_out_uncond=None if out_u is None else out_u.to(d), cfg=cfg, t_next=tn.to(d),
                                alphas_cumprod=sch.alphas_cumprod.to(d), next_dtype=lowp, want_eps=True)
    assert torch.equal(e.cpu(), eps)
    assert torch.equal(x0.cpu(), x0_ref)
    assert torch.equal(xn.cpu(), xn_ref)


def test_solver_step_bad_prediction_type():
    from tdm_b200 import ops
    with pytest.raises(ValueError, match="currently not supported"):
        ops.solver_step(torch.zeros(1, 8, 2, 2, device="cuda"), torch.zeros(1, 4, 2, 2, device="cuda"),
                        torch.zeros(1, dtype=torch.long, device="cuda"), torch.ones(10, device="cuda"),
                        torch.ones(10, device="cuda"), "sample")


@pytest.mark.parametrize("lowp", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("mixed", [False, True])
def test_renoise_bit_exact(lowp, mixed):
    from tdm_b200 import ops
    torch.manual_seed(1)
    B, C, H, W = 4, 4, 8, 8
    sch, a, s = _tables(lowp)
    pred = glue.OraclePredictor(sch, a, s, None, None, lowp=lowp)
    x = torch.randn(B, C, H, W).to(lowp)
    nz = torch.randn(B, C, H, W).to(lowp)
    t1 = torch.tensor([225, 450, 675, 10])
    t2 = torch.tensor([300, 450, 800, 999])  # t2 >= t1 (valid domain)
    ref = pred.obtain_mixed_noise(x, nz, t1, t2) if mixed else pred.add_noise(x, nz, t1, t2)
    d = "cuda"
    out = ops.renoise(x.to(d), nz.to(d), t1.to(d), t2.to(d), a.to(d), s.to(d), ref.dtype, mixed=mixed)
    assert torch.equal(out.cpu(), ref)


def test_renoise_invalid_interval_is_nan_like_reference():
    from tdm_b200 import ops
    sch, a, s = _tables(torch.float16)
    pred = glue.OraclePredictor(sch, a, s, None, None)
    x = torch.ones(1, 4, 2, 2, dtype=torch.float16)
    t1, t2 = torch.tensor([675]), torch.tensor([300])  # t2 < t1 -> beta^2 < 0 -> NaN (D3)
    ref = pred.add_noise(x, x, t1, t2)
    out = ops.renoise(x.cuda(), x.cuda(), t1.cuda(), t2.cuda(), a.cuda(), s.cuda(), torch.float16)
    assert torch.isnan(ref).all() and torch.isnan(out).all()


def test_select_and_renoise_matches_gather():
    """main.py:466,473-479 with the per-sample gather of SURVEY.md D2; index selection is bit-exact."""
    from tdm_b200 import ops
    torch.manual_seed(2)
    B, C, H, W = 6, 4, 8, 8
    lowp = torch.float16
    sch, a, s = _tables(lowp)
    pred = glue.OraclePredictor(sch, a, s, None, None)
    states = [torch.randn(B, C, H, W)] + [torch.randn(B, C, H, W).to(lowp) for _ in range(4)]  # reversed list
    ind_t = torch.randint(1, 5, (B,))
    t1 = ind_t * 900 // 4 - 1  # the states live at 224/449/674/899
    t2 = torch.minimum(t1 + torch.randint(0, 100, (B,)), torch.tensor(999))
    nz = torch.randn(B, C, H, W).to(lowp)
    gathered = torch.stack([st.to(lowp) for st in states])[ind_t, torch.arange(B)]
    ref = pred.add_noise(gathered, nz, t1, t2)
    d = "cuda"
    out = ops.renoise([st.to(d) for st in states], nz.to(d), t1.to(d), t2.to(d), a.to(d), s.to(d), lowp,
                      index=ind_t.to(d))
    assert torch.equal(out.cpu(), ref)


@pytest.mark.parametrize("use_huber", [True, False])
@pytest.mark.parametrize("cfg", [4.5, 1.0])
def test_loss_and_gradient(use_huber, cfg):
    from tdm_b200 import ops
    torch.manual_seed(3)
    B, C, H, W = 2, 4, 32, 32
    m = torch.randn(B, C, H, W, requires_grad=True)
    sd = torch.randn(B, C, H, W)
    su = torch.randn(B, C, H, W)
    ref = glue.tdm_loss(m, sd, su, cfg, use_huber)
    ref.backward()
    loss, grad = ops.tdm_loss(m.detach().cuda(), sd.cuda(), su.cuda(), cfg, use_huber)
    assert loss.dtype == torch.float64  # main.py:519-524: fp32 / fp64 promotion
    # The same restatement evaluated by torch eager on the GPU (IEEE sqrt): only the fp64 summation
    # order differs -> 1e-12.  torch's CPU sqrt goes through MKL VML (0.55 ulp, measured here), so
    # against the CPU oracle the Huber branch agrees to ~1e-9; the bound is 1e-8 relative.
    ref_gpu = glue.tdm_loss(m.detach().cuda(), sd.cuda(), su.cuda(), cfg, use_huber)
    assert abs(loss.item() - ref_gpu.item()) <= 1e-12 * abs(ref_gpu.item())
    assert abs(loss.item() - ref.item()) <= 1e-8 * abs(ref.item())
    # gradient: closed form evaluated in fp64 then rounded to fp32; autograd works in fp32/fp64 mixed.
    # tolerance 2e-6 relative to the largest gradient entry (a few fp32 ulps).
    g_ref = m.grad
    assert (grad.cpu() - g_ref).abs().max().item() <= 2e-6 * g_ref.abs().max().item()
    g_cf = glue.tdm_loss_grad(m.detach(), sd, su, cfg, use_huber)
    assert (grad.cpu().double() - g_cf).abs().max().item() <= 2e-7 * g_cf.abs().max().item()
