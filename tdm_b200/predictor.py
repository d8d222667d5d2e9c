"""Mirror of the reference's src/predictor.py `Predictor` (same constructor, methods, argument meaning,
return arity -- including the THREE values of `return_double=True`, predictor.py:59-60) on the library's
fused kernels.  `predict` is differentiable w.r.t. the score model's parameters (main.py:512 calls it on
the student with grad enabled): the fused solver step is an autograd.Function whose backward is the
library's tdm_solver_step_bwd kernel.
"""
from __future__ import annotations

import torch

from . import _lib, ops
from ._lib import check, ptr, stream_handle
from . import models as _models


class _SolverStepFn(torch.autograd.Function):
    """(eps_cfg, x0) = fused CFG combine + predicted_origin; gradient flows to the model outputs only."""

    @staticmethod
    def forward(ctx, out_c, out_u, x_t, t, alphas, sigmas, cfg, pred_type):
        eps, x0, _ = ops.solver_step(out_c, x_t, t, alphas, sigmas, pred_type, model_out_uncond=out_u,
                                     cfg=cfg if out_u is not None else 1.0, want_eps=True)
        ctx.save_for_backward(t, alphas, sigmas)
        ctx.meta = (cfg, out_u is not None, out_c.shape, x_t.shape[1], pred_type)
        return eps, x0

    @staticmethod
    def backward(ctx, g_eps, g_x0):
        t, alphas, sigmas = ctx.saved_tensors
        cfg, has_u, shape, C_, pred_type = ctx.meta
        if pred_type != "epsilon":
            raise _lib.TdmError("backward of the fused solver step is implemented for epsilon prediction")
        B, Cout = shape[0], shape[1]
        HW = 1
        for s in shape[2:]:
            HW *= s
        d_c = torch.empty(shape, device=t.device, dtype=torch.float32)
        d_u = torch.empty(shape, device=t.device, dtype=torch.float32) if has_u else None
        g_eps = g_eps.contiguous().float() if g_eps is not None else None
        g_x0 = g_x0.contiguous().float() if g_x0 is not None else None
        check(_lib.lib().tdm_solver_step_bwd(ptr(g_eps), ptr(g_x0), float(cfg), int(has_u), ptr(t), ptr(alphas),
                                             ptr(sigmas), _lib.dtype_enum(alphas), alphas.numel(), ptr(d_c),
                                             ptr(d_u), B, C_, Cout, HW, stream_handle()))
        return d_c, d_u, None, None, None, None, None, None


def _step(out_c, out_u, x_t, t, alphas, sigmas, cfg, pred_type):
    if torch.is_grad_enabled() and (out_c.requires_grad or (out_u is not None and out_u.requires_grad)):
        return _SolverStepFn.apply(out_c, out_u, x_t, t, alphas, sigmas, cfg, pred_type)
    eps, x0, _ = ops.solver_step(out_c, x_t, t, alphas, sigmas, pred_type, model_out_uncond=out_u,
                                 cfg=cfg if out_u is not None else 1.0, want_eps=True)
    return eps, x0


class Predictor:
    def __init__(self, noise_scheduler, alpha_schedule, sigma_schedule, uncond_prompt_embeds, uncond_attention_mask,
                 lowp=None):
        self.noise_scheduler = noise_scheduler
        self.alpha_schedule = alpha_schedule
        self.sigma_schedule = sigma_schedule
        self.uncond_prompt_embeds = uncond_prompt_embeds
        self.uncond_attention_mask = uncond_attention_mask
        self.added_cond_kwargs = {"resolution": None, "aspect_ratio": None}
        self.lowp = lowp  # None -> models.LOWP_DTYPE at call time (reference: torch.float16, predictor.py:85)

    def _uncond(self, B):
        """The reference builds the unconditional context with batch == train_batch_size (main.py:417-423) and
        breaks for any other B; a single row is broadcast here."""
        c, m = self.uncond_prompt_embeds, self.uncond_attention_mask
        if c.shape[0] != B:
            c, m = c[:1].expand(B, *c.shape[1:]), m[:1].expand(B, *m.shape[1:])
        return c, m

    # The reference evaluates the conditional and the unconditional branch of classifier-free guidance as two model
    # calls (predictor.py:33-39).  No op of the DiT couples samples, so under no_grad both branches run here as ONE
    # call on the stacked batch [cond; uncond] (2B rows per GEMM: better tile/wave efficiency, half the launches);
    # the per-sample results are bit-identical to the two separate calls (tests/test_iteration_gpu.py).
    batch_cfg = True

    def _predict_pair(self, score_model, x, t, ehs, mask):
        if not hasattr(score_model, "_forward_impl") or ehs is None or mask is None:
            return None
        B = x.shape[0]
        uc, um = self._uncond(B)
        if uc.shape[1:] != ehs.shape[1:] or um.shape[1:] != mask.shape[1:]:
            return None
        t = t.expand(B) if t.dim() == 0 else t
        x2 = torch.cat([x, x])
        t2 = torch.cat([t, t])
        e2 = torch.cat([ehs, uc.to(ehs.dtype)])
        m2 = torch.cat([mask, um.to(mask.dtype)])
        out = score_model(x2, timestep=t2, added_cond_kwargs=self.added_cond_kwargs, encoder_hidden_states=e2,
                          encoder_attention_mask=m2, return_dict=False)[0]
        return out[:B], out[B:]

    def predict(self, score_model, noisy_samples, timesteps, encoder_hidden_states, prompt_attention_mask, cfg=None,
                steps=1, return_double=False, timestep_cond=None, return_uncond=False):
        """predictor.py:30-74.  `return_uncond` (extension, default off) appends the unconditional x0 estimate to
        the `return_double` tuple: the quantity main.py:508 names `sd_latents_uncond`."""
        pred_type = self.noise_scheduler.config.prediction_type
        for _ in range(steps):
            t_long = timesteps.long()
            pair = None
            if cfg is not None and self.batch_cfg and not torch.is_grad_enabled():
                pair = self._predict_pair(score_model, noisy_samples, timesteps, encoder_hidden_states, prompt_attention_mask)
            if pair is not None:
                out_c, out_u = pair
            else:
                out_c = score_model(noisy_samples, timestep=timesteps, added_cond_kwargs=self.added_cond_kwargs,
                                    encoder_hidden_states=encoder_hidden_states,
                                    encoder_attention_mask=prompt_attention_mask, return_dict=False)[0]
            if cfg is not None:
                if pair is None:
                    uc, um = self._uncond(noisy_samples.shape[0])
                    out_u = score_model(noisy_samples, timestep=timesteps, added_cond_kwargs=self.added_cond_kwargs,
                                        encoder_hidden_states=uc, encoder_attention_mask=um, return_dict=False)[0]
                score_pred_cfg, pred_latents = _step(out_c, out_u, noisy_samples, t_long, self.alpha_schedule,
                                                     self.sigma_schedule, cfg, pred_type)       # predictor.py:42-49
                score_pred, pred_latents_nocfg = _step(out_c, None, noisy_samples, t_long, self.alpha_schedule,
                                                       self.sigma_schedule, 1.0, pred_type)      # predictor.py:50-56
                timesteps = timesteps - timesteps // steps
                if return_double:
                    if return_uncond:
                        _, pred_latents_uncond = _step(out_u, None, noisy_samples, t_long, self.alpha_schedule,
                                                       self.sigma_schedule, 1.0, pred_type)
                        return score_pred_cfg, pred_latents, pred_latents_nocfg, pred_latents_uncond
                    return score_pred_cfg, pred_latents, pred_latents_nocfg
                # predictor.py:58.  With steps == 1 the re-noised sample is never used and is skipped.
                # Deliberate deviation for steps > 1 under autograd: the reference detaches pred_latents only, so a
                # gradient could reach the next step's INPUT through score_pred_cfg; the scheduler kernel is not
                # differentiable and the hot path (main.py:507-512, steps=1) never takes that route, so both operands
                # are detached here.  Values are identical.
                if steps > 1:
                    noisy_samples = self.noise_scheduler.add_noise(pred_latents.detach(), score_pred_cfg.detach(),
                                                                   timesteps.long())
            else:
                score_pred, pred_latents = _step(out_c, None, noisy_samples, t_long, self.alpha_schedule,
                                                 self.sigma_schedule, 1.0, pred_type)
                timesteps = timesteps - timesteps // steps
                if steps > 1:
                    noisy_samples = self.noise_scheduler.add_noise(pred_latents.detach(), score_pred.detach(), timesteps)
        if cfg is not None:
            return score_pred_cfg, pred_latents
        return score_pred, pred_latents

    def add_noise(self, samples, noise, t1, t2):
        """predictor.py:76-85."""
        return ops.renoise(samples, noise, t1.long(), t2.long(), self.alpha_schedule, self.sigma_schedule,
                           self.lowp or _models.LOWP_DTYPE)

    def select_and_add_noise(self, states, index, noise, t1, t2):
        """main.py:479 with the per-sample gather of SURVEY.md D2 fused into the same kernel:
        sample b starts from states[index[b]][b]."""
        return ops.renoise(list(states), noise, t1.long(), t2.long(), self.alpha_schedule, self.sigma_schedule,
                           self.lowp or _models.LOWP_DTYPE, index=index.long())

    def obtain_mixed_noise(self, model_noise, noise, t1, t2):
        """predictor.py:87-97."""
        out_dtype = torch.promote_types(torch.promote_types(model_noise.dtype, self.alpha_schedule.dtype), noise.dtype)
        return ops.renoise(model_noise, noise, t1.long(), t2.long(), self.alpha_schedule, self.sigma_schedule,
                           out_dtype, mixed=True)
