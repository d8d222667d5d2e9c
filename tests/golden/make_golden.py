"""Generate tests/golden/reference_glue.pt from the REAL reference (run where /root/reference exists).

    python tests/golden/make_golden.py

What is executed is the reference's own code, unmodified:
  * src/utils.py, src/models.py, src/predictor.py imported through oracle/reference_import.py;
  * the loss block src/main.py:514-529 and the index/timestep draws src/main.py:473-478, which live
    inside main() and cannot be imported (main.py needs accelerate/diffusers), are exec()'d from the
    file's own text, line range asserted, in a namespace that supplies their free variables.
The score model is tests/helpers.FakeScoreModel (IEEE add/mul only) and the scheduler is the oracle's
DDPMScheduler restatement (diffusers is not installable here), so the fixture pins the reference-owned
glue, not the third-party transformer.
"""
from __future__ import annotations

import sys
import textwrap
import types
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from helpers import FakeScoreModel, synthetic_inputs  # noqa: E402
from oracle import reference_import  # noqa: E402
from oracle.schedule import OracleDDPMScheduler  # noqa: E402

MAIN_PY = reference_import.REFERENCE_ROOT / "src" / "main.py"


def reference_lines(first: int, last: int, must_start_with: str) -> str:
    lines = MAIN_PY.read_text().splitlines()[first - 1:last]
    assert lines[0].strip().startswith(must_start_with), (first, lines[0])
    return textwrap.dedent("\n".join(lines))


def run_reference_loss(model_latents, sd_latents, sd_latents_uncond, cfg, use_huber):
    src = reference_lines(514, 529, "if args.cfg > 1:")
    ns = dict(torch=torch, F=F, args=types.SimpleNamespace(cfg=cfg, use_huber=use_huber),
              model_latents=model_latents, sd_latents=sd_latents, sd_latents_uncond=sd_latents_uncond)
    exec(src, ns)
    return ns["loss"], ns["weighting_factor"]


def run_reference_draws(seed, bsz, total_steps, use_separate):
    """main.py:473-478 on the CPU generator: ind_t, t_fake, noise_fake in the reference's draw order."""
    src = reference_lines(473, 478, "ind_t = torch.randint(1, 5")
    torch.manual_seed(seed)
    noise = torch.randn([bsz, 4, 8, 8]).to(torch.float16)  # main.py:443
    ns = dict(torch=torch, bsz=bsz, noise=noise, fw_t=240, total_steps=total_steps,
              args=types.SimpleNamespace(use_separate=use_separate))
    exec(src, ns)
    return noise, ns["ind_t"], ns["t_fake"], ns["noise_fake"]


def main():
    utils, models, predictor_mod = reference_import.load()
    g = {}
    lowp = torch.float16
    sch = OracleDDPMScheduler()
    alphas = torch.sqrt(sch.alphas_cumprod).to(lowp)
    sigmas = torch.sqrt(1 - sch.alphas_cumprod).to(lowp)
    noise, ctx, mask, uctx, umask = synthetic_inputs()
    g["inputs"] = dict(noise=noise, ctx=ctx, mask=mask, uctx=uctx, umask=umask)

    # utils.py
    t = torch.tensor([899, 224])
    eps = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(11))
    g["extract"] = dict(t=t, out=utils.extract_into_tensor(alphas, t, noise.shape))
    g["predicted_origin"] = dict(
        eps=eps, t=t,
        x0_eps=utils.predicted_origin(eps, t, noise, "epsilon", alphas, sigmas),
        x0_v=utils.predicted_origin(eps, t, noise, "v_prediction", alphas, sigmas))

    # models.py generate_new: training call (main.py:465), sampling call with mid_points (README.md:129), CFG
    for name, kw in {
        "train_4step": dict(steps=4, return_mid=True, total_steps=900),
        "sample_midpoints": dict(steps=4, return_mid=True, mid_points=[999, 856, 665, 399, 0], total_steps=1000),
        "sample_1step": dict(steps=1, total_steps=900),
        "sample_cfg": dict(steps=2, return_mid=True, total_steps=900,
                           add_cfg={"uncond_attention_mask": umask, "uncond_prompt_embeds": uctx, "cfg": 7.5}),
    }.items():
        model = FakeScoreModel()
        torch.manual_seed(5)
        out = models.generate_new(model, sch, noise, noise, ctx, mask, **kw)
        rec = dict(kwargs={k: v for k, v in kw.items() if k != "add_cfg"}, has_cfg="add_cfg" in kw,
                   timesteps=torch.stack(model.calls))
        if kw.get("return_mid"):
            rec["imgs_list"], rec["noisy_imgs_list"] = out
        else:
            rec["latent"] = out
        g["generate_new/" + name] = rec

    # predictor.py
    P = predictor_mod.Predictor(sch, alphas, sigmas, uctx, umask)
    x = noise
    tt = torch.tensor([300, 440])
    for name, kw in {"cfg": dict(cfg=4.5), "nocfg": dict(cfg=None), "double": dict(cfg=4.5, return_double=True),
                     "cfg_2steps": dict(cfg=4.5, steps=2)}.items():
        out = P.predict(FakeScoreModel(), x, tt, ctx, mask, **kw)
        g["predict/" + name] = dict(t=tt, kwargs=kw, out=list(out))
    t1 = torch.tensor([225, 450])
    t2 = torch.tensor([300, 450])
    nz = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(12)).to(lowp)
    g["add_noise"] = dict(t1=t1, t2=t2, noise=nz, out=P.add_noise(x, nz, t1, t2))
    g["mixed_noise"] = dict(t1=t1, t2=t2, noise=nz, model_noise=eps, out=P.obtain_mixed_noise(eps, nz, t1, t2))

    # main.py:514-529 loss, all four branches
    gen = torch.Generator().manual_seed(13)
    m, sd, su = (torch.randn(2, 4, 8, 8, generator=gen) for _ in range(3))
    g["loss_inputs"] = dict(model_latents=m, sd_latents=sd, sd_latents_uncond=su)
    for cfg in (4.5, 1.0):
        for huber in (True, False):
            mm = m.clone().requires_grad_(True)
            loss, w = run_reference_loss(mm, sd, su, cfg, huber)
            loss.backward()
            g[f"loss/cfg{cfg}_huber{int(huber)}"] = dict(cfg=cfg, use_huber=huber, loss=loss.detach(), weight=w,
                                                         grad=mm.grad.clone())

    # main.py:473-478 draw order
    for sep in (True, False):
        noise_d, ind_t, t_fake, noise_fake = run_reference_draws(21, 8, 900, sep)
        g[f"draws/separate{int(sep)}"] = dict(seed=21, bsz=8, total_steps=900, ind_t=ind_t, t_fake=t_fake,
                                              noise=noise_d, noise_fake=noise_fake)

    out = Path(__file__).with_name("reference_glue.pt")
    torch.save(g, out)
    print("wrote", out, out.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
