"""Mirror of the reference's src/models.py: `generate_new`, the K-step deterministic sampler (x0-predict ->
re-noise loop, DDIM == first-order DPM-solver) used both for the student trajectory (main.py:465) and for
4-NFE sampling (training.py:111).  Same signature and return structure; per step ONE fused kernel does
eps-half select + CFG combine + predicted_origin + scheduler.add_noise + low-precision cast
(models.py:40-58), and the DiT evaluations run in the sm_100a kernels of Transformer2DModel.
"""
from __future__ import annotations

import torch

from . import ops

# The reference hard-codes torch.float16 at its cast site (models.py:58); north_star asks for bf16.
# The cast SITE is kept, the dtype is a module-level knob (DESIGN.md "dtype policy").
LOWP_DTYPE = torch.float16

_ADDED = {"resolution": None, "aspect_ratio": None}


def _call(unet, x, t, ctx, mask):
    return unet(x, timestep=t, added_cond_kwargs=_ADDED, encoder_hidden_states=ctx, encoder_attention_mask=mask,
                return_dict=False)[0]


def generate_new(unet, noise_scheduler, latent, noise, encoder_hidden_states, prompt_attention_mask,
                 steps=1, return_mid=False, mid_points=None, total_steps=800, add_cfg=None, lowp=None):
    lowp = lowp or LOWP_DTYPE
    if add_cfg is not None:
        uncond_attention_mask = add_cfg["uncond_attention_mask"]
        uncond_prompt_embeds = add_cfg["uncond_prompt_embeds"]
        cfg = add_cfg["cfg"]
    dev = latent.device
    # same draw as models.py:26 (constant value, but it advances the device generator)
    T_ = torch.randint(total_steps - 1, total_steps, (latent.shape[0],), device=dev).long()
    zero_t = torch.zeros_like(T_)
    acp = noise_scheduler.alphas_cumprod_on(dev) if hasattr(noise_scheduler, "alphas_cumprod_on") \
        else noise_scheduler.alphas_cumprod.to(dev).float().contiguous()
    alpha_schedule = torch.sqrt(noise_scheduler.alphas_cumprod).to(dev).to(latent.dtype)       # models.py:29
    sigma_schedule = torch.sqrt(1 - noise_scheduler.alphas_cumprod).to(dev).to(latent.dtype)   # models.py:30
    imgs_list, noisy_imgs_list = [], []
    pure_noisy = noise
    pred_type = noise_scheduler.config.prediction_type
    for ind in range(steps):
        noisy_imgs_list.append(pure_noisy)
        out_c = _call(unet, pure_noisy, T_, encoder_hidden_states, prompt_attention_mask)
        out_u = None
        if add_cfg is not None:
            out_u = _call(unet, pure_noisy, T_, uncond_prompt_embeds, uncond_attention_mask)
        T_next = (mid_points[ind + 1] + zero_t) if mid_points is not None else (T_ - total_steps // steps)
        _, latent, pure_noisy = ops.solver_step(
            out_c, pure_noisy, T_, alpha_schedule, sigma_schedule, pred_type, model_out_uncond=out_u,
            cfg=cfg if add_cfg is not None else 1.0, t_next=T_next, alphas_cumprod=acp, next_dtype=lowp)
        imgs_list.append(latent)
        T_ = T_next
    noisy_imgs_list.append(latent)
    if return_mid:
        return imgs_list, noisy_imgs_list
    return latent
