"""ORACLE (test infrastructure, never on the product path): the TDM training iteration in plain PyTorch
(fp32 weights, autograd, torch.optim.AdamW, torch.nn.utils.clip_grad_norm_), CPU-runnable.

Two restatements of src/main.py:441-540:

* `literal_draws` -- the RNG / index draws exactly as the reference makes them (main.py:443-478).
* `RepairedIteration.step` -- the iteration north_star describes, built from the reference's own functions
  as restated in oracle/glue.py (generate_new, Predictor.add_noise / predict, predicted_origin, the loss of
  main.py:514-529) with the defects of SURVEY.md 3.5 repaired:
    D2  `noisy_imgs_list[ind_t]` is a per-sample gather;
    D3  k = ind_t - 1 selects one of the four student-produced states, t1 = k*total_steps//4,
        t2 = t1 + (t_fake_draw - 240) (clamped to 999) with --use_separate, else spread over [t1, 1000);
    D4  predict(..., return_double=True) returns 3 values; sd_latents = cond x0, sd_latents_uncond = uncond x0;
    D5  the fake-score model is trained: MSE between its epsilon prediction at (latents_fake, t2) and the
        effective noise (latents_fake - alpha_t2 * x0_sel) / sigma_t2, own AdamW;
    D6  the student may carry LoRA adapters (oracle/lora.py);
    D7  teacher and fake model are evaluated on the noised sample latents_fake; the loss of main.py:514-529
        is taken with model_latents := the fake model's x0 estimate (value kept in the graph of the student
        through the usual stop-gradient identity): d loss / d x0_student = d loss / d model_latents, where
        x0_student is the student's x0 prediction of the trajectory step that produced the selected state.
  The same definition is implemented on the GPU by tdm_b200/training.py; tests/test_iteration_gpu.py
  compares the two.  PARITY UNPINNED against the reference for this function: the reference's main() cannot
  execute one iteration, so only its constituent functions are pinned (tests/golden).
"""
from __future__ import annotations

import torch

from . import glue

FW_T = 240


def literal_draws(bsz, latent_hw, total_steps, use_separate, lowp=torch.float16, device="cpu"):
    """main.py:443, 452, 455, 473-478 in order (generate_new's own draw, models.py:26, happens in between)."""
    h, w = latent_hw
    noise = torch.randn([bsz, 4, h, w]).to(lowp).to(device)                              # :443
    torch.randn_like(noise)                                                               # :452 (unused)
    torch.randint(total_steps - 1, total_steps, (bsz,), device=noise.device).long()      # :455 (constant)
    return noise


def fake_branch_draws(noise, total_steps, use_separate):
    bsz = noise.shape[0]
    ind_t = torch.randint(1, 5, (bsz,), device=noise.device).long()                      # :473
    if use_separate:
        t_fake = torch.randint(FW_T, FW_T + 200, (bsz,), device=noise.device).long()     # :475
    else:
        t_fake = torch.randint(FW_T, total_steps, (bsz,), device=noise.device).long()    # :477
    noise_fake = torch.randn_like(noise)                                                  # :478
    return ind_t, t_fake, noise_fake


def repaired_levels(ind_t, t_fake_draw, total_steps, use_separate, num_train_timesteps=1000, steps=4):
    k = ind_t - 1
    t1 = k * total_steps // steps
    off = t_fake_draw - FW_T
    t2 = t1 + off if use_separate else t1 + off * (num_train_timesteps - t1) // (total_steps - FW_T)
    t2 = torch.clamp(t2, max=num_train_timesteps - 1)
    t_in = (total_steps - 1) - (steps - 1 - k) * (total_steps // steps)
    return k, t1, t2, t_in


def _gather(states, idx):
    """D2: out[b] = states[idx[b]][b]."""
    stacked = torch.stack([s.float() for s in states])
    return stacked[idx, torch.arange(idx.shape[0], device=idx.device)]


_KW = {"resolution": None, "aspect_ratio": None}


class RepairedIteration:
    def __init__(self, student, fake, teacher, scheduler, uncond_ctx, uncond_mask, *, cfg=4.5, total_steps=900,
                 use_huber=True, use_separate=True, lowp=torch.float16, lr=2e-5, lr_fake=2e-5, max_grad_norm=1.0,
                 weight_decay=1e-2):
        self.student, self.fake, self.teacher = student, fake, teacher
        self.scheduler = scheduler
        self.cfg, self.total_steps, self.use_huber, self.use_separate, self.lowp = cfg, total_steps, use_huber, use_separate, lowp
        self.alphas = torch.sqrt(scheduler.alphas_cumprod).to(lowp)          # main.py:136-139
        self.sigmas = torch.sqrt(1 - scheduler.alphas_cumprod).to(lowp)
        self.predictor = glue.OraclePredictor(scheduler, self.alphas, self.sigmas, uncond_ctx, uncond_mask, lowp=lowp)
        for p in teacher.parameters():
            p.requires_grad_(False)
        self.max_grad_norm = max_grad_norm
        sp = [p for p in student.parameters() if p.requires_grad]
        fp = [p for p in fake.parameters() if p.requires_grad]
        self.opt_student = torch.optim.AdamW(sp, lr=lr, betas=(0.9, 0.999), weight_decay=weight_decay, eps=1e-8)
        self.opt_fake = torch.optim.AdamW(fp, lr=lr_fake, betas=(0.9, 0.999), weight_decay=weight_decay, eps=1e-8)
        self.last = {}

    def _eps(self, model, x, t, ctx, mask):
        return model(x, timestep=t, added_cond_kwargs=_KW, encoder_hidden_states=ctx, encoder_attention_mask=mask,
                     return_dict=False)[0].chunk(2, dim=1)[0]

    def step(self, noise, ctx, mask, draws=None, apply=True):
        B = noise.shape[0]
        torch.randn_like(noise)                                                                     # main.py:452
        torch.randint(self.total_steps - 1, self.total_steps, (B,), device=noise.device)            # main.py:455
        with torch.no_grad():                                                                       # main.py:464-466
            imgs_list, states = glue.few_step_sample(self.student, self.scheduler, noise, noise, ctx, mask, steps=4,
                                                     return_mid=True, total_steps=self.total_steps, lowp=self.lowp)
        states.reverse()
        if draws is None:
            draws = fake_branch_draws(noise, self.total_steps, self.use_separate)
        ind_t, t_fake_draw, noise_fake = draws
        k, t1, t2, t_in = repaired_levels(ind_t, t_fake_draw, self.total_steps, self.use_separate)
        # per-sample select fused with predictor.add_noise: the kernel reads each state in its own dtype
        sel = torch.stack([self.predictor.add_noise(s, noise_fake, t1, t2) for s in states[:4]])
        latents_fake = sel[k, torch.arange(B, device=k.device)]                                                      # main.py:479 (D2)
        x0_sel = _gather(imgs_list[::-1], k)
        x_in = _gather(states[1:5], k).to(self.lowp)
        pt = self.scheduler.config.prediction_type
        # fake-score model
        eps_fake = self._eps(self.fake, latents_fake, t2, ctx, mask)
        x0_fake = glue.x0_from_model_output(eps_fake.detach(), t2, latents_fake, pt, self.alphas, self.sigmas)
        eps_target = glue.x0_from_model_output(x0_sel, t2, latents_fake, "epsilon", self.sigmas, self.alphas)
        loss_fake = ((eps_fake.double() - eps_target.double()) ** 2).mean()
        # teacher (D4 mapping)
        with torch.no_grad():
            eps_c = self._eps(self.teacher, latents_fake, t2, ctx, mask)
            eps_u = self._eps(self.teacher, latents_fake, t2, self.predictor.uncond_ctx.expand(B, -1, -1),
                              self.predictor.uncond_mask.expand(B, -1))
            sd = glue.x0_from_model_output(eps_c, t2, latents_fake, pt, self.alphas, self.sigmas)
            sdu = glue.x0_from_model_output(eps_u, t2, latents_fake, pt, self.alphas, self.sigmas)
        m = x0_fake.clone().requires_grad_(True)
        loss = glue.tdm_loss(m, sd, sdu, self.cfg, self.use_huber)
        (g_x0,) = torch.autograd.grad(loss, m)
        # student step that produced the selected state, with grad
        eps_s = self._eps(self.student, x_in, t_in, ctx, mask)
        x0_s = glue.x0_from_model_output(eps_s, t_in, x_in, pt, self.alphas, self.sigmas)
        self.opt_student.zero_grad(set_to_none=True)
        self.opt_fake.zero_grad(set_to_none=True)
        x0_s.backward(g_x0.float())
        loss_fake.backward()
        self.last = dict(ind_t=ind_t, t_fake=t2, t1=t1, t_in=t_in, latents_fake=latents_fake, x0_fake=x0_fake,
                         eps_target=eps_target, sd=sd, sdu=sdu, g_x0=g_x0, x_in=x_in, x0_sel=x0_sel)
        if apply:
            for opt in (self.opt_student, self.opt_fake):
                params = [p for g in opt.param_groups for p in g["params"]]
                if self.max_grad_norm:
                    torch.nn.utils.clip_grad_norm_(params, self.max_grad_norm)
                opt.step()
        return {"loss": loss.detach(), "loss_fake": loss_fake.detach()}
