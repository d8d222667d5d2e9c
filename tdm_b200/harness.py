"""Offline harness: what the reference's setup (main.py:132-139, 161-176, 367-375, 416-433) provides, without the
network -- random-init PixArt-style DiTs (student / fake score / teacher start as identical copies, as all three
come from the same checkpoint in the reference), the DDPM schedule, synthetic T5 embeddings and masks
(SURVEY.md 8d "Synthetic inputs").  Host-side plumbing only."""
from __future__ import annotations

import torch

from .scheduler import DDPMScheduler
from .transformer import DEFAULT_CONFIG, TINY_CONFIG, Transformer2DModel

CONFIGS = {
    "tiny": dict(TINY_CONFIG),                          # BASELINE.json configs[0]
    "pixart-512": dict(DEFAULT_CONFIG),                 # configs[1..3]
    "pixart-1024": dict(DEFAULT_CONFIG, sample_size=128),  # configs[4]
}


def make_scheduler():
    """main.py:132-134."""
    s = DDPMScheduler(beta_start=0.0001, beta_end=0.02, beta_schedule="linear", steps_offset=1, clip_sample=False)
    s.config.prediction_type = "epsilon"
    return s


def make_models(config="pixart-512", seed=0, device="cuda", lora_rank=64, lora_alpha=None, teacher=True, fake=True):
    cfg = CONFIGS[config] if isinstance(config, str) else dict(config)
    student = Transformer2DModel(cfg, device=device, seed=seed)
    fake_m = teacher_m = None
    if fake:
        fake_m = Transformer2DModel(cfg, device=device, init_weights=False)
        fake_m.copy_weights_from(student)
        fake_m.requires_grad_(True).train()
    if teacher:
        teacher_m = Transformer2DModel(cfg, device=device, init_weights=False)
        teacher_m.copy_weights_from(student)
        teacher_m.requires_grad_(False).eval()
    if lora_rank:
        student.add_lora(rank=lora_rank, alpha=lora_alpha, seed=seed + 1)
    return student, fake_m, teacher_m


def synthetic_text(B, L=120, Cc=4096, seed=0, pin=False):
    """ctx ~ N(0,1) [B,L,Cc]; mask = arange(L) < len_b, len_b ~ U{1..L}; one unconditional row whose mask keeps
    only token 0 (an empty prompt tokenises to EOS alone, main.py:417-423).  HOST tensors."""
    g = torch.Generator().manual_seed(seed)
    ctx = torch.randn(B, L, Cc, generator=g)
    lens = torch.randint(1, L + 1, (B,), generator=g)
    mask = (torch.arange(L)[None] < lens[:, None]).to(torch.int64)
    uctx = torch.randn(1, L, Cc, generator=g)
    umask = torch.zeros(1, L, dtype=torch.int64)
    umask[:, 0] = 1
    out = [ctx, mask, uctx, umask]
    if pin and torch.cuda.is_available():
        out = [t.pin_memory() for t in out]
    return tuple(out)
