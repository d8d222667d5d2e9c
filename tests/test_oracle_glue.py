"""CPU tests that pin the oracle (oracle/glue.py, oracle/schedule.py):
  (1) against the committed golden vectors produced by the REAL reference (tests/golden/make_golden.py);
  (2) against the real reference functions imported live, wherever /root/reference exists;
  (3) known answers from SURVEY.md 8c (visited timesteps, schedule constants).
Integer results (timesteps, indices) and every elementwise result are compared bit-exactly."""
from pathlib import Path

import pytest
import torch

from helpers import FakeScoreModel, synthetic_inputs
from oracle import glue, reference_import
from oracle.schedule import OracleDDPMScheduler

GOLD = torch.load(Path(__file__).parent / "golden" / "reference_glue.pt", weights_only=False)
LOWP = torch.float16


def _setup():
    sch = OracleDDPMScheduler()
    a = torch.sqrt(sch.alphas_cumprod).to(LOWP)
    s = torch.sqrt(1 - sch.alphas_cumprod).to(LOWP)
    i = GOLD["inputs"]
    return sch, a, s, i["noise"], i["ctx"], i["mask"], i["uctx"], i["umask"]


def test_schedule_known_answers():
    sch = OracleDDPMScheduler()
    ac = sch.alphas_cumprod
    assert ac.shape == (1000,) and ac.dtype == torch.float32
    assert abs(ac[899].sqrt().item() - 0.016589) < 5e-7
    assert abs((1 - ac[899]).sqrt().item() - 0.999862) < 5e-7
    assert abs(ac[999].sqrt().item() - 0.006353) < 5e-7
    assert sch.config.prediction_type == "epsilon"
    sch.register_to_config(prediction_type="v_prediction")
    assert sch.config.prediction_type == "v_prediction"


def test_golden_utils():
    sch, a, s, noise, *_ = _setup()
    g = GOLD["extract"]
    assert torch.equal(glue.table_lookup(a, g["t"], noise.ndim), g["out"])
    g = GOLD["predicted_origin"]
    assert torch.equal(glue.x0_from_model_output(g["eps"], g["t"], noise, "epsilon", a, s), g["x0_eps"])
    assert torch.equal(glue.x0_from_model_output(g["eps"], g["t"], noise, "v_prediction", a, s), g["x0_v"])
    with pytest.raises(ValueError, match="currently not supported"):
        glue.x0_from_model_output(g["eps"], g["t"], noise, "sample", a, s)


@pytest.mark.parametrize("name", ["train_4step", "sample_midpoints", "sample_1step", "sample_cfg"])
def test_golden_generate_new(name):
    sch, a, s, noise, ctx, mask, uctx, umask = _setup()
    g = GOLD["generate_new/" + name]
    kw = dict(g["kwargs"])
    if g["has_cfg"]:
        kw["add_cfg"] = {"uncond_attention_mask": umask, "uncond_prompt_embeds": uctx, "cfg": 7.5}
    model = FakeScoreModel()
    torch.manual_seed(5)
    out = glue.few_step_sample(model, sch, noise, noise, ctx, mask, **kw)
    assert torch.equal(torch.stack(model.calls), g["timesteps"])  # visited timesteps, bit-exact
    if "latent" in g:
        assert torch.equal(out, g["latent"])
    else:
        imgs, states = out
        assert len(imgs) == len(g["imgs_list"]) and len(states) == len(g["noisy_imgs_list"])
        for x, y in zip(imgs + states, g["imgs_list"] + g["noisy_imgs_list"]):
            assert x.dtype == y.dtype and torch.equal(x, y)


def test_visited_timesteps_known_answers():
    """(steps 4, total 900) visits 899/674/449/224; mid_points visits README.md:129's list."""
    assert GOLD["generate_new/train_4step"]["timesteps"][:, 0].tolist() == [899, 674, 449, 224]
    assert GOLD["generate_new/sample_midpoints"]["timesteps"][:, 0].tolist() == [999, 856, 665, 399]
    g = GOLD["generate_new/train_4step"]
    assert len(g["imgs_list"]) == 4 and len(g["noisy_imgs_list"]) == 5
    assert g["noisy_imgs_list"][1].dtype == torch.float16 and g["imgs_list"][0].dtype == torch.float32


@pytest.mark.parametrize("name", ["cfg", "nocfg", "double", "cfg_2steps"])
def test_golden_predict(name):
    sch, a, s, noise, ctx, mask, uctx, umask = _setup()
    g = GOLD["predict/" + name]
    P = glue.OraclePredictor(sch, a, s, uctx, umask)
    out = P.predict(FakeScoreModel(), noise, g["t"], ctx, mask, **g["kwargs"])
    assert len(out) == len(g["out"])  # return_double returns THREE values
    for x, y in zip(out, g["out"]):
        assert x.dtype == y.dtype and torch.equal(x, y)


def test_golden_renoise():
    sch, a, s, noise, ctx, mask, uctx, umask = _setup()
    P = glue.OraclePredictor(sch, a, s, uctx, umask)
    g = GOLD["add_noise"]
    out = P.add_noise(noise, g["noise"], g["t1"], g["t2"])
    assert out.dtype == torch.float16 and torch.equal(out, g["out"])
    g = GOLD["mixed_noise"]
    assert torch.equal(P.obtain_mixed_noise(g["model_noise"], g["noise"], g["t1"], g["t2"]), g["out"])


@pytest.mark.parametrize("cfg", [4.5, 1.0])
@pytest.mark.parametrize("huber", [True, False])
def test_golden_loss_and_closed_form_gradient(cfg, huber):
    i = GOLD["loss_inputs"]
    g = GOLD[f"loss/cfg{cfg}_huber{int(huber)}"]
    m = i["model_latents"].clone().requires_grad_(True)
    loss = glue.tdm_loss(m, i["sd_latents"], i["sd_latents_uncond"], cfg, huber)
    assert loss.dtype == torch.float64 == g["loss"].dtype
    # identical op sequence; 1e-8 leaves room for a different CPU's vectorised sqrt (MKL VML, 0.55 ulp)
    assert abs(loss.item() - g["loss"].item()) <= 1e-8 * abs(g["loss"].item())
    loss.backward()
    assert (m.grad - g["grad"]).abs().max() <= 1e-6 * g["grad"].abs().max()
    cf = glue.tdm_loss_grad(i["model_latents"], i["sd_latents"], i["sd_latents_uncond"], cfg, huber)
    assert (cf - g["grad"].double()).abs().max() <= 1e-6 * g["grad"].abs().max()  # closed form == reference autograd


@pytest.mark.parametrize("sep", [True, False])
def test_golden_draw_order(sep):
    """main.py:443, 473-478: randn(noise) -> randint(ind_t) -> randint(t_fake) -> randn_like(noise_fake)."""
    from tdm_b200.training import draw_fake_branch  # host-side mirror, torch generator only
    g = GOLD[f"draws/separate{int(sep)}"]
    torch.manual_seed(g["seed"])
    noise = torch.randn([g["bsz"], 4, 8, 8]).to(torch.float16)
    ind_t, t_fake, noise_fake = draw_fake_branch(noise, g["total_steps"], use_separate=sep)
    assert torch.equal(noise, g["noise"])
    assert torch.equal(ind_t, g["ind_t"]) and ind_t.dtype == torch.int64
    assert torch.equal(t_fake, g["t_fake"]) and t_fake.dtype == torch.int64
    assert torch.equal(noise_fake, g["noise_fake"])
    assert int(ind_t.min()) >= 1 and int(ind_t.max()) <= 4
    assert int(t_fake.min()) >= 240 and int(t_fake.max()) < (440 if sep else 900)


needs_ref = pytest.mark.skipif(not reference_import.available(), reason="/root/reference not on this machine")


@needs_ref
def test_live_reference_random_sweep():
    """Oracle vs the live reference on fresh random inputs (more shapes / timesteps than the fixture)."""
    utils, models, predictor_mod = reference_import.load()
    sch = OracleDDPMScheduler()
    for lowp in (torch.float16, torch.float32):
        a = torch.sqrt(sch.alphas_cumprod).to(lowp)
        s = torch.sqrt(1 - sch.alphas_cumprod).to(lowp)
        for seed in range(3):
            noise, ctx, mask, uctx, umask = synthetic_inputs(B=3, H=4, W=6, seed=seed, lowp=lowp)
            Pr = predictor_mod.Predictor(sch, a, s, uctx, umask)
            Po = glue.OraclePredictor(sch, a, s, uctx, umask)
            g = torch.Generator().manual_seed(seed)
            t1 = torch.randint(0, 900, (3,), generator=g)
            t2 = torch.minimum(t1 + torch.randint(0, 300, (3,), generator=g), torch.tensor(999))
            nz = torch.randn(noise.shape, generator=g).to(lowp)
            assert torch.equal(Pr.add_noise(noise, nz, t1, t2), Po.add_noise(noise, nz, t1, t2))
            assert torch.equal(Pr.obtain_mixed_noise(noise.float(), nz, t1, t2),
                               Po.obtain_mixed_noise(noise.float(), nz, t1, t2))
            for kw in (dict(cfg=3.0), dict(cfg=None), dict(cfg=2.0, return_double=True)):
                ro = Pr.predict(FakeScoreModel(), noise, t2, ctx, mask, **kw)
                oo = Po.predict(FakeScoreModel(), noise, t2, ctx, mask, **kw)
                assert all(torch.equal(x, y) for x, y in zip(ro, oo)) and len(ro) == len(oo)
            if lowp == torch.float16:  # generate_new hard-codes the fp16 cast (models.py:58)
                torch.manual_seed(seed)
                r = models.generate_new(FakeScoreModel(), sch, noise, noise, ctx, mask, steps=3, return_mid=True,
                                        total_steps=900)
                torch.manual_seed(seed)
                o = glue.few_step_sample(FakeScoreModel(), sch, noise, noise, ctx, mask, steps=3, return_mid=True,
                                         total_steps=900)
                assert all(torch.equal(x, y) for x, y in zip(r[0] + r[1], o[0] + o[1]))


@needs_ref
def test_reference_defects_reproduced():
    """D3 (NaN when t2 < t1) and D4 (three return values) are properties of the reference."""
    _, _, predictor_mod = reference_import.load()
    sch, a, s, noise, ctx, mask, uctx, umask = _setup()
    P = predictor_mod.Predictor(sch, a, s, uctx, umask)
    assert torch.isnan(P.add_noise(noise, noise, torch.tensor([675, 900]), torch.tensor([300, 300]))).all()
    assert len(P.predict(FakeScoreModel(), noise, torch.tensor([300, 300]), ctx, mask, cfg=4.5, return_double=True)) == 3
