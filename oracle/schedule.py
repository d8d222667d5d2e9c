"""ORACLE (test infrastructure): restatement of the parts of `diffusers.DDPMScheduler` the reference
touches (main.py:132-139, 449; models.py:29-30, 49, 58; predictor.py:47, 58, 70).

Third-party, un-vendored (diffusers >= 0.26.0.dev0, main.py:68).  Known answers used to pin it
(SURVEY.md 8c): alpha_899 = 0.016589, sigma_899 = 0.999862, alpha_999 = 0.006353.
"""
from __future__ import annotations

import torch


class _Config:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class OracleDDPMScheduler:
    def __init__(self, beta_start=0.0001, beta_end=0.02, beta_schedule="linear", num_train_timesteps=1000,
                 steps_offset=1, trained_betas=None, clip_sample=False, rescale_betas_zero_snr=False,
                 prediction_type="epsilon"):
        if beta_schedule != "linear" or trained_betas is not None or rescale_betas_zero_snr:
            raise NotImplementedError("the reference only builds the linear schedule (main.py:132-133)")
        self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.config = _Config(prediction_type=prediction_type, num_train_timesteps=num_train_timesteps,
                              beta_start=beta_start, beta_end=beta_end, beta_schedule=beta_schedule,
                              steps_offset=steps_offset, clip_sample=clip_sample)

    def register_to_config(self, **kw):
        self.config.__dict__.update(kw)

    def add_noise(self, original_samples, noise, timesteps):
        # diffusers moves alphas_cumprod to the sample's device AND dtype, then takes **0.5 there.
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        timesteps = timesteps.to(original_samples.device)
        a = ac[timesteps] ** 0.5
        s = (1 - ac[timesteps]) ** 0.5
        a = a.flatten()
        s = s.flatten()
        while a.ndim < original_samples.ndim:
            a = a.unsqueeze(-1)
            s = s.unsqueeze(-1)
        return a * original_samples + s * noise
