"""Host side of the TDM distillation iteration (mirror of the reference's src/main.py:441-540 and of the
4-NFE sampling call in src/training.py:107-112), driving the library's sm_100a kernels.

Two things live here, kept apart on purpose:

* the reference's own RNG / index draws, restated literally (`draw_iteration_noise`, `draw_fake_branch`):
  they stay `torch.randint` / `torch.randn` calls on torch's generator, in the reference's order, so
  timesteps and trajectory indices are bit-identical to the reference for the same seed;
* the REPAIRED iteration (`TdmTrainer.step`): the schedule north_star asks for -- frozen teacher forward,
  fake-score forward+backward, LoRA student forward+backward on the 4-step trajectory, score-difference
  loss -- built from the reference's functions (`generate_new`, `Predictor.*`, `predicted_origin`, the loss
  of main.py:514-529) with the defects D2-D7 repaired as DESIGN.md lists them.

No arithmetic happens in PyTorch here: tensors are allocated by torch and every value is produced by a
library kernel.  There is no CPU fallback.
"""
from __future__ import annotations

import torch

from . import ops
from . import models as _models
from .models import generate_new
from .predictor import Predictor

FW_T = 240  # main.py:468


# ----------------------------------------------------------------------------------------------------
# literal draws (main.py:443, 452, 455, 473-478)
# ----------------------------------------------------------------------------------------------------
def draw_iteration_noise(bsz, latent_hw, lowp, device):
    """main.py:443: noise drawn on the CPU generator, cast to low precision, moved to the device."""
    h, w = latent_hw
    return torch.randn([bsz, 4, h, w]).to(lowp).to(device)


def draw_fake_branch(noise, total_steps, use_separate, fw_t=FW_T):
    """main.py:473-478: ind_t, t_fake, noise_fake in the reference's draw order, on noise.device's generator."""
    bsz = noise.shape[0]
    ind_t = torch.randint(1, 5, (bsz,), device=noise.device).long()
    if use_separate:
        t_fake = torch.randint(fw_t, fw_t + 200, (bsz,), device=noise.device).long()
    else:
        t_fake = torch.randint(fw_t, total_steps, (bsz,), device=noise.device).long()
    noise_fake = torch.randn_like(noise)
    return ind_t, t_fake, noise_fake


def repaired_levels(ind_t, t_fake_draw, total_steps, use_separate, fw_t=FW_T, num_train_timesteps=1000, steps=4):
    """Integer arithmetic that turns the literal draws into a valid (state, t1, t2) triple (repair of D3).

    k = ind_t - 1 selects the k-th entry of the REVERSED trajectory [x0, x@t_1, .., x@t_3] -- the four states
    the student produced (index 4 of the reference's list is the input noise, which the student cannot
    influence); t1 = k * total_steps // steps is the reference's own level formula (main.py:479) applied to k;
    t2 >= t1 re-uses the drawn offset: t ~ [t_k, t_k + 200) with --use_separate, t ~ [t_k, T) otherwise
    (README.md:180-183).  int64, left-to-right precedence as in the reference."""
    k = ind_t - 1
    t1 = k * total_steps // steps
    off = t_fake_draw - fw_t
    if use_separate:
        t2 = t1 + off
    else:
        t2 = t1 + off * (num_train_timesteps - t1) // (total_steps - fw_t)
    t2 = torch.clamp(t2, max=num_train_timesteps - 1)
    # student step that produced state k: forward step j = steps-1-k, evaluated at T_j (models.py:26, 57)
    t_in = (total_steps - 1) - (steps - 1 - k) * (total_steps // steps)
    return k, t1, t2, t_in


# ----------------------------------------------------------------------------------------------------
# clip + AdamW on the flat arenas (main.py:206-224, 536-540)
# ----------------------------------------------------------------------------------------------------
class FlatAdamW:
    """torch.optim.AdamW semantics (lr, betas, eps, weight_decay as args.py:204-213) over a model's flat fp32
    arenas, one fused kernel per arena: clip coefficient applied on the fly, bf16 weight copy refreshed."""

    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=1.0):
        self.model = model
        self.lr, self.betas, self.eps, self.wd, self.max_grad_norm = lr, betas, eps, weight_decay, max_grad_norm
        self.t = 0
        model._sync_bf16()
        model._ensure_grads()
        self.groups = []  # (param, grad, bf16)
        if model.train_base:
            self.groups.append((model._flat, model._grad, model._flat16))
        if model._lflat is not None:
            self.groups.append((model._lflat, model._lgrad, model._lflat16))
        dev = model.device
        self.state = [(torch.zeros_like(p), torch.zeros_like(p)) for p, _, _ in self.groups]
        self._sumsq = torch.zeros((), device=dev, dtype=torch.float64)
        self._coef = torch.ones((), device=dev, dtype=torch.float32)
        self.grad_norm = torch.zeros((), device=dev, dtype=torch.float32)

    def step(self, grad_world_scale=1.0, lr=None):
        """`grad_world_scale`: factor the (summed over ranks) gradients still carry, e.g. 1/world_size."""
        self.t += 1
        for i, (_, g, _) in enumerate(self.groups):
            ops.sumsq(g, self._sumsq, zero_first=(i == 0))
        ops.clip_coef(self._sumsq, self.max_grad_norm if self.max_grad_norm else 0.0, grad_world_scale, self._coef,
                      self.grad_norm)
        for (p, g, p16), (m, v) in zip(self.groups, self.state):
            ops.adamw_step(p, g, m, v, p16, lr if lr is not None else self.lr, self.betas[0], self.betas[1], self.eps,
                           self.wd, self.t, self._coef)
        self.model.mark_updated()

    def zero_grad(self):
        for _, g, _ in self.groups:
            g.zero_()


class EMAModel:
    """diffusers.training_utils.EMAModel as the reference uses it (main.py:188-189 ctor, 325-326 .to(device),
    553-554 .step(unet.parameters())): an exponential moving average of the student's trainable parameters, kept as
    one flat fp32 shadow per arena and updated by one kernel launch.  Decay schedule of EMAModel without warm-up:
    min(decay, (1 + step) / (10 + step))."""

    def __init__(self, model, decay=0.9999, min_decay=0.0):
        self.model, self.decay, self.min_decay = model, decay, min_decay
        self.optimization_step = 0
        self.shadow = [p.detach().clone() for p in model.flat_params()]

    def get_decay(self, optimization_step):
        step = max(0, optimization_step - 1)
        if step <= 0:
            return 0.0
        return max(self.min_decay, min(self.decay, (1 + step) / (10 + step)))

    def step(self, parameters=None):
        from ._lib import check, ptr, stream_handle, lib
        self.optimization_step += 1
        d = self.get_decay(self.optimization_step)
        for sh, p in zip(self.shadow, self.model.flat_params()):
            check(lib().tdm_ema_update(ptr(sh), ptr(p), p.numel(), float(d), stream_handle()))

    def copy_to(self, model):
        with torch.no_grad():
            for sh, p in zip(self.shadow, model.flat_params()):
                p.copy_(sh)
        model.mark_updated()
        model._flat16_version = model._lflat16_version = -1

    def to(self, device=None, dtype=None):
        self.shadow = [s.to(device=device) for s in self.shadow]
        return self


# ----------------------------------------------------------------------------------------------------
# the repaired iteration
# ----------------------------------------------------------------------------------------------------
class TdmTrainer:
    """One TDM distillation iteration on one GPU (data-parallel ranks each own one; `grad_sync` couples them).

    student / fake / teacher: tdm_b200.transformer.Transformer2DModel (student usually with add_lora()).
    Per iteration and per sample: 4 student forwards (trajectory, no grad) + 1 student forward/backward,
    1 fake-score forward/backward, 2 teacher forwards (cond + uncond)."""

    def __init__(self, student, fake, teacher, scheduler, uncond_ctx, uncond_mask, *, cfg=4.5, total_steps=900,
                 use_huber=True, use_separate=True, lowp=torch.bfloat16, lr=2e-5, lr_fake=2e-5, max_grad_norm=1.0,
                 weight_decay=1e-2, grad_sync=None, overlap_streams=None, defer_fake_opt=True):
        self.student, self.fake, self.teacher = student, fake, teacher
        self.scheduler = scheduler
        self.cfg, self.total_steps, self.use_huber, self.use_separate = cfg, total_steps, use_huber, use_separate
        self.lowp = lowp
        dev = student.device
        self.device = dev
        # main.py:136-139: sqrt tables moved to the device and cast to the low-precision dtype
        self.alpha_schedule = torch.sqrt(scheduler.alphas_cumprod).to(dev).to(lowp)
        self.sigma_schedule = torch.sqrt(1 - scheduler.alphas_cumprod).to(dev).to(lowp)
        self.predictor = Predictor(scheduler, self.alpha_schedule, self.sigma_schedule, uncond_ctx, uncond_mask, lowp=lowp)
        teacher.requires_grad_(False)
        self.opt_student = FlatAdamW(student, lr=lr, weight_decay=weight_decay, max_grad_norm=max_grad_norm)
        self.opt_fake = FlatAdamW(fake, lr=lr_fake, weight_decay=weight_decay, max_grad_norm=max_grad_norm)
        self.grad_sync = grad_sync
        self.world = grad_sync.world_size if grad_sync is not None else 1
        self.last = {}
        # Two independent chains follow the trajectory: (fake-score forward -> its backward) and (teacher forward ->
        # student forward).  On two streams the second chain's kernels fill the first chain's wave tails and the SMs its
        # small LoRA-sized launches leave idle; they join at the loss.  Results are unchanged (same kernels, same order
        # per tensor).  TDM_B200_OVERLAP=0 runs everything on one stream.
        import os as _os
        if overlap_streams is None:
            overlap_streams = _os.environ.get("TDM_B200_OVERLAP", "1") != "0"
        self._side = torch.cuda.Stream(device=dev) if overlap_streams else None
        # Optional: the fake-score optimiser step (HBM-bound, ~2.4 ms) of iteration i runs on the side stream under the
        # trajectory forwards of iteration i+1 (which only touch the student); the fake-score forward waits for it.
        # Readers of the fake-score arenas from outside (checkpoint.py, tests) call sync() first; bench.py closes its
        # timed regions with sync() so exactly K optimiser steps are inside.  On by default (TDM_B200_DEFER_OPT=0 or
        # defer_fake_opt=False restores the in-order step).
        if _os.environ.get("TDM_B200_DEFER_OPT", "1") == "0":
            defer_fake_opt = False
        self.defer_fake_opt = bool(defer_fake_opt) and self._side is not None
        self._fake_opt_evt = None

    # -- pieces -------------------------------------------------------------------------------------
    def _trajectory(self, noise, ctx, mask):
        """main.py:464-466."""
        with torch.no_grad():
            imgs_list, noisy_imgs_list = generate_new(self.student, self.scheduler, noise, noise, ctx, mask, steps=4,
                                                      return_mid=True, total_steps=self.total_steps, lowp=self.lowp)
        noisy_imgs_list.reverse()
        return imgs_list, noisy_imgs_list

    def step(self, noise, ctx, mask, draws=None):
        """noise: [B,4,H,W] low-precision device tensor (main.py:443); ctx [B,L,Cc]; mask [B,L].
        Returns {'loss': fp64 0-dim, 'loss_fake': fp64 0-dim} device tensors (no host sync)."""
        S, Fk, T = self.student, self.fake, self.teacher
        pr = self.predictor
        B = noise.shape[0]
        if ctx.dtype != torch.bfloat16:  # one cast per iteration instead of one per DiT evaluation
            ctx = ops.cast_bf16(ctx if ctx.dtype == torch.float32 else ctx.float())
        # main.py:452, 455: draws whose values are unused but which advance the device generator
        torch.randn_like(noise)
        torch.randint(self.total_steps - 1, self.total_steps, (B,), device=noise.device)
        imgs_list, rev_states = self._trajectory(noise, ctx, mask)
        if draws is None:
            draws = draw_fake_branch(noise, self.total_steps, self.use_separate)
        ind_t, t_fake_draw, noise_fake = draws
        k, t1, t_fake, t_in = repaired_levels(ind_t, t_fake_draw, self.total_steps, self.use_separate)
        # D2: per-sample gather of the selected state, fused with the t1 -> t2 re-noising (predictor.py:76-85)
        latents_fake = pr.select_and_add_noise(rev_states[:4], k, noise_fake, t1, t_fake)
        x0_sel = ops.select_states(imgs_list[::-1], k)                      # the clean estimate behind that state
        x_in = ops.select_states(rev_states[1:5], k, out_dtype=self.lowp)   # the state the producing step started from
        side = self._side
        main = torch.cuda.current_stream()
        gs = self.grad_sync
        from ._lib import check, ptr, stream_handle, lib, dtype_enum
        if side is not None:
            side.wait_stream(main)  # latents_fake, x_in, ctx are ready; the teacher / student were last used on `main`
        # ---- side chain, part 1: frozen teacher (cond + uncond, x0 space; main.py:507-510, D4 mapping), then the
        #      student forward with activations kept (its inputs do not depend on the loss) ----
        with torch.cuda.stream(side if side is not None else main):
            with torch.no_grad():
                _, _, sd_latents, sd_latents_uncond = pr.predict(T, latents_fake, t_fake, ctx, mask, cfg=self.cfg, steps=1,
                                                                 return_double=True, return_uncond=True)
            out_s, saved_s = S._forward_impl(x_in, t_in, ctx, mask, True)
        # ---- main chain: fake-score model forward (saved) ----
        self.sync()  # a deferred optimiser step of the previous iteration has to land first
        out_fake, saved_f = Fk._forward_impl(latents_fake, t_fake, ctx, mask, True)
        _, x0_fake, _ = ops.solver_step(out_fake, latents_fake, t_fake, self.alpha_schedule, self.sigma_schedule,
                                        self.scheduler.config.prediction_type)   # main.py:484-490
        # effective noise of latents_fake w.r.t. x0_sel: (x - alpha x0) / sigma == predicted_origin with the tables swapped
        _, eps_target, _ = ops.solver_step(x0_sel, latents_fake, t_fake, self.sigma_schedule, self.alpha_schedule, "epsilon")
        loss_fake, g_fake = ops.mse_loss(out_fake, eps_target)
        x0_ready = torch.cuda.Event()
        x0_ready.record(main)
        # ---- main chain: fake-score backward.  Issued BEFORE the student's backward: the exchange runs in order on the
        #      process group's one communication stream, and the fake-score gradients are final first (their backward
        #      needs only the fake-score loss, the student's waits for the teacher and for x0_fake).  Issued after it, the
        #      16 fake-score buckets (2.4 GB) queued behind the student's last bucket and ran un-overlapped at the end of
        #      the iteration (profiles/r02_dp_timeline_2gpu_before.json: 5.7 ms exposed at 2 GPUs). ----
        Fk.backward_impl(saved_f, g_fake, gs.hook(Fk) if gs is not None else None)
        if gs is not None:
            gs.finish(Fk)
        # ---- side chain, part 2: score-difference loss (main.py:514-529) on the fake-vs-teacher x0 estimates; its
        #      closed-form gradient lands on the student's x0 prediction of the step that produced the state; student
        #      backward through the solver step ----
        with torch.cuda.stream(side if side is not None else main):
            if side is not None:
                side.wait_event(x0_ready)
            loss, g_x0 = ops.tdm_loss(x0_fake, sd_latents, sd_latents_uncond, self.cfg, self.use_huber)
            d_out_s = torch.empty_like(out_s)
            check(lib().tdm_solver_step_bwd(0, ptr(g_x0), 1.0, 0, ptr(t_in), ptr(self.alpha_schedule),
                                            ptr(self.sigma_schedule), dtype_enum(self.alpha_schedule),
                                            self.alpha_schedule.numel(), ptr(d_out_s), 0, B, noise.shape[1], out_s.shape[1],
                                            noise[0, 0].numel(), stream_handle()))
            # data parallel: every rank issues its bucketed all-reduces in the same host order (fake-score buckets, then
            # student buckets); each one waits for the stream it was issued from
            S.backward_impl(saved_s, d_out_s, gs.hook(S) if gs is not None else None)
            if gs is not None:
                gs.finish(S)
        if side is not None:
            main.wait_stream(side)
        if gs is not None:
            gs.wait()
        inv = 1.0 / self.world
        self.opt_student.step(inv)
        self.opt_student.zero_grad()
        if self.defer_fake_opt:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                self.opt_fake.step(inv)
                self.opt_fake.zero_grad()
                self._fake_opt_evt = torch.cuda.Event()
                self._fake_opt_evt.record(side)
        else:
            self.opt_fake.step(inv)
            self.opt_fake.zero_grad()
        # references only (small [B,4,H,W] tensors): what the parity tests compare with the oracle's intermediates
        self.last = {"loss": loss, "loss_fake": loss_fake, "ind_t": ind_t, "t_fake": t_fake, "t1": t1, "t_in": t_in,
                     "latents_fake": latents_fake, "x0_fake": x0_fake, "eps_target": eps_target, "sd": sd_latents,
                     "sdu": sd_latents_uncond, "g_x0": g_x0, "x_in": x_in, "x0_sel": x0_sel}
        return {"loss": loss, "loss_fake": loss_fake}

    def sync(self):
        """Make the current stream wait for a deferred fake-score optimiser step (no-op otherwise)."""
        if self._fake_opt_evt is not None:
            torch.cuda.current_stream().wait_event(self._fake_opt_evt)
            self._fake_opt_evt = None

    def step_from_host(self, noise_host, ctx_host, mask_host):
        """Public end-to-end entry: pinned HOST buffers in, host loss out (what bench.py's `e2e` times)."""
        dev = self.device
        noise = noise_host.to(dev, non_blocking=True)
        ctx = ctx_host.to(dev, non_blocking=True)
        mask = mask_host.to(dev, non_blocking=True)
        out = self.step(noise, ctx, mask)
        both = torch.stack([out["loss"], out["loss_fake"]]).to("cpu")  # device -> host read of the step's result
        return float(both[0]), float(both[1])


# ----------------------------------------------------------------------------------------------------
# 4-NFE sampling (training.py:107-112; README.md:129 timesteps)
# ----------------------------------------------------------------------------------------------------
SAMPLING_MID_POINTS = [999, 856, 665, 399, 0]


def sample_4nfe(student, scheduler, noise, ctx, mask, lowp=None):
    """generate_new(..., steps=4, mid_points=[999,856,665,399,0], total_steps=1000): final x0 latents."""
    with torch.no_grad():
        return generate_new(student, scheduler, noise, noise, ctx, mask, steps=4, mid_points=SAMPLING_MID_POINTS,
                            total_steps=1000, lowp=lowp or _models.LOWP_DTYPE)


def save_validation_images(unet, noise_scheduler, vae, fixed_c, fixed_mask, fixed_noise, total_steps, global_step=None,
                           output_dir=None, lowp=None):
    """training.py:107-121: the 4-step and the 1-step student samples of the fixed prompts / noise, decoded by TAESD,
    clamped and mapped to [0, 1].  Returns (images_4step, images_1step) fp32 [B,3,8H,8W]; written as JPEG grids like the
    reference (fixed_4step_<step>.jpg, fixed_1step_<step>.jpg, nrow=2) when `output_dir` is given and torchvision is
    importable (JPEG encoding is host-side I/O, outside the hot path)."""
    from .taesd import decode_to_images
    lowp = lowp or _models.LOWP_DTYPE
    with torch.no_grad():
        lat4 = generate_new(unet, noise_scheduler, fixed_noise, fixed_noise, fixed_c, fixed_mask, steps=4,
                            total_steps=total_steps, lowp=lowp)
        lat1 = generate_new(unet, noise_scheduler, fixed_noise, fixed_noise, fixed_c, fixed_mask, steps=1,
                            total_steps=total_steps, lowp=lowp)
        img4 = decode_to_images(vae, lat4)
        img1 = decode_to_images(vae, lat1)
    if output_dir is not None:
        try:
            from torchvision.utils import save_image
            save_image(img1[:4], f"{output_dir}/fixed_1step_{global_step}.jpg", normalize=False, nrow=2)
            save_image(img4[:4], f"{output_dir}/fixed_4step_{global_step}.jpg", normalize=False, nrow=2)
        except ImportError:
            torch.save({"fixed_4step": img4[:4].cpu(), "fixed_1step": img1[:4].cpu()}, f"{output_dir}/fixed_{global_step}.pt")
    return img4, img1


def sample_4nfe_from_host(student, scheduler, noise_host, ctx_host, mask_host, lowp=None):
    dev = student.device
    lat = sample_4nfe(student, scheduler, noise_host.to(dev, non_blocking=True), ctx_host.to(dev, non_blocking=True),
                      mask_host.to(dev, non_blocking=True), lowp)
    return lat.to("cpu")
