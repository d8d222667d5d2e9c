"""ORACLE (test infrastructure, never on the product path): CPU restatement of the reference's own
solver / schedule / loss glue, function by function.  Each function cites the reference lines it
follows.  It is pinned (tests/test_oracle_glue.py, run wherever /root/reference exists) against
the REAL reference functions imported unmodified through oracle/reference_import.py, and against the
committed golden vectors in tests/golden/ (generated from the real reference by
tests/golden/make_golden.py).

`lowp` is the dtype the reference hard-codes as torch.float16 at its cast sites (models.py:58,
predictor.py:85, main.py:138-139, 443); north_star asks for bf16, so it is a parameter here and the
default keeps the literal fp16.
"""
from __future__ import annotations

import torch


def table_lookup(table: torch.Tensor, t: torch.Tensor, ndim: int) -> torch.Tensor:
    """utils.py:70-73 extract_into_tensor: gather schedule entries and shape them [B,1,1,...]."""
    picked = torch.gather(table, -1, t)
    return picked.reshape(t.shape[0], *([1] * (ndim - 1)))


def x0_from_model_output(pred, t, x_t, prediction_type, alphas, sigmas):
    """utils.py:47-59 predicted_origin."""
    if prediction_type not in ("epsilon", "v_prediction"):
        raise ValueError(f"Prediction type {prediction_type} currently not supported.")
    s = table_lookup(sigmas, t, x_t.ndim)
    a = table_lookup(alphas, t, x_t.ndim)
    if prediction_type == "epsilon":
        return (x_t - s * pred) / a
    return a * x_t - s * pred


_COND_KW = {"resolution": None, "aspect_ratio": None}


def _eps(model, x, t, ctx, mask):
    """One score-model evaluation keeping the epsilon half (models.py:38-40, predictor.py:33-35)."""
    out = model(x, timestep=t, added_cond_kwargs=_COND_KW, encoder_hidden_states=ctx,
                encoder_attention_mask=mask, return_dict=False)[0]
    return out.chunk(2, dim=1)[0]


def few_step_sample(model, scheduler, latent, noise, ctx, mask, steps=1, return_mid=False, mid_points=None,
                    total_steps=800, add_cfg=None, lowp=torch.float16):
    """models.py:20-62 generate_new: K-step x0-predict -> re-noise loop (DDIM == DPM-solver-1)."""
    B = latent.shape[0]
    t = torch.randint(total_steps - 1, total_steps, (B,), device=latent.device).long()  # models.py:26-27
    zeros = torch.zeros_like(t)
    alphas = torch.sqrt(scheduler.alphas_cumprod).to(latent.device).to(latent.dtype)      # models.py:29
    sigmas = torch.sqrt(1 - scheduler.alphas_cumprod).to(latent.device).to(latent.dtype)  # models.py:30
    x = noise
    x0_list, state_list = [], []
    x0 = latent
    for k in range(steps):
        state_list.append(x)
        eps = _eps(model, x, t, ctx, mask)
        if add_cfg is not None:                                                              # models.py:41-45
            eps_u = _eps(model, x, t, add_cfg["uncond_prompt_embeds"], add_cfg["uncond_attention_mask"])
            eps = eps_u + add_cfg["cfg"] * (eps - eps_u)
        x0 = x0_from_model_output(eps, t, x, scheduler.config.prediction_type, alphas, sigmas)
        x0_list.append(x0)
        t = (mid_points[k + 1] + zeros) if mid_points is not None else (t - total_steps // steps)
        x = scheduler.add_noise(x0, eps, t).to(lowp)                                        # models.py:58
    state_list.append(x0)
    return (x0_list, state_list) if return_mid else x0


class OraclePredictor:
    """predictor.py:20-97."""

    def __init__(self, scheduler, alphas, sigmas, uncond_ctx, uncond_mask, lowp=torch.float16):
        self.scheduler, self.alphas, self.sigmas = scheduler, alphas, sigmas
        self.uncond_ctx, self.uncond_mask = uncond_ctx, uncond_mask
        self.lowp = lowp

    def predict(self, model, x, t, ctx, mask, cfg=None, steps=1, return_double=False):
        """predictor.py:30-74."""
        kind = lambda: self.scheduler.config.prediction_type
        for _ in range(steps):
            eps = _eps(model, x, t, ctx, mask)
            if cfg is not None:
                eps_u = _eps(model, x, t, self.uncond_ctx, self.uncond_mask)
                eps_cfg = eps_u + cfg * (eps - eps_u)                                        # predictor.py:42
                x0 = x0_from_model_output(eps_cfg, t.long(), x, kind(), self.alphas, self.sigmas)
                x0_nocfg = x0_from_model_output(eps, t.long(), x, kind(), self.alphas, self.sigmas)
                t = t - t // steps
                x = self.scheduler.add_noise(x0.detach(), eps_cfg, t.long())
                if return_double:
                    return eps_cfg, x0, x0_nocfg                                             # predictor.py:59-60
            else:
                x0 = x0_from_model_output(eps, t, x, kind(), self.alphas, self.sigmas)
                t = t - t // steps
                x = self.scheduler.add_noise(x0.detach(), eps, t)
        return (eps_cfg, x0) if cfg is not None else (eps, x0)

    def _coeffs(self, t1, t2, ndim):
        s1 = table_lookup(self.sigmas, t1, ndim)
        a1 = table_lookup(self.alphas, t1, ndim)
        s2 = table_lookup(self.sigmas, t2, ndim)
        a2 = table_lookup(self.alphas, t2, ndim)
        beta = (s2 ** 2 - (a2 / a1 * s1) ** 2) ** 0.5                                       # predictor.py:82-83
        return a1, a2, s2, beta

    def add_noise(self, x, noise, t1, t2):
        """predictor.py:76-85: move a sample from noise level t1 to t2 >= t1 (NaN when t2 < t1, D3)."""
        a1, a2, _, beta = self._coeffs(t1, t2, x.ndim)
        return (x / a1 * a2 + beta * noise).to(self.lowp)

    def obtain_mixed_noise(self, eps, noise, t1, t2):
        """predictor.py:87-97 (never called by the reference; epsilon-space target of a t1->t2 move)."""
        a1, a2, s2, beta = self._coeffs(t1, t2, eps.ndim)
        return (eps / a1 * a2 + beta * noise) / s2


def tdm_loss(model_latents, sd_latents, sd_latents_uncond, cfg, use_huber, huber_c=1e-3):
    """main.py:514-529.  Returns the fp64 scalar loss."""
    coop = sd_latents + (cfg - 1) * (sd_latents - sd_latents_uncond) if cfg > 1 else sd_latents
    sd_cfg = sd_latents + (cfg - 1) * (sd_latents - sd_latents_uncond)                      # main.py:518
    w = torch.abs(model_latents.double() - sd_cfg.double()).mean(dim=[1, 2, 3], keepdim=True).detach()
    diff = model_latents.float() - coop.detach().float()
    if use_huber:
        return torch.mean((torch.sqrt(diff ** 2 + huber_c ** 2) - huber_c) / w)
    return (diff ** 2 / w).mean()


def tdm_loss_grad(model_latents, sd_latents, sd_latents_uncond, cfg, use_huber, huber_c=1e-3):
    """Closed-form dL/d(model_latents) of tdm_loss (checked against autograd in
    tests/test_oracle_glue.py).  The weighting factor is detached in the reference (main.py:519)."""
    coop = sd_latents + (cfg - 1) * (sd_latents - sd_latents_uncond) if cfg > 1 else sd_latents
    sd_cfg = sd_latents + (cfg - 1) * (sd_latents - sd_latents_uncond)
    w = torch.abs(model_latents.double() - sd_cfg.double()).mean(dim=[1, 2, 3], keepdim=True)
    diff = model_latents.float() - coop.float()
    n = model_latents.numel()
    if use_huber:
        return diff / torch.sqrt(diff ** 2 + huber_c ** 2) / (w * n)
    return 2 * diff / (w * n)
