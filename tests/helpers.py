"""Shared test fixtures: a deterministic stand-in score model and synthetic inputs.

The fake model uses only IEEE add/mul on fp32 (no reductions, no transcendentals), so its output is
bit-reproducible on any CPU or GPU; it returns the [B, 2C, H, W] fp32 tensor shape the reference
expects from `unet(...)[0]` (models.py:38-40)."""
from __future__ import annotations

import torch


class FakeScoreModel:
    def __init__(self, gain=0.7):
        self.gain = gain
        self.calls = []

    def __call__(self, x, timestep=None, added_cond_kwargs=None, encoder_hidden_states=None,
                 encoder_attention_mask=None, return_dict=False):
        assert added_cond_kwargs == {"resolution": None, "aspect_ratio": None}
        self.calls.append(timestep.clone())
        xf = x.float()
        t = (timestep.float() * 0.001).view(-1, 1, 1, 1)
        c = encoder_hidden_states.float()[:, 0, 0].view(-1, 1, 1, 1)
        m = encoder_attention_mask.float()[:, -1].view(-1, 1, 1, 1)
        eps = xf * self.gain + (xf * xf) * t * 0.1 + c * 0.25 - m * 0.125
        return (torch.cat([eps, eps * 0.5 + 1.0], dim=1),)


def synthetic_inputs(B=2, C=4, H=8, W=8, L=6, Cc=16, seed=0, lowp=torch.float16, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    noise = torch.randn(B, C, H, W, generator=g).to(lowp)
    ctx = torch.randn(B, L, Cc, generator=g)
    lens = torch.randint(1, L + 1, (B,), generator=g)
    mask = (torch.arange(L)[None] < lens[:, None]).to(torch.int64)
    uctx = torch.randn(1, L, Cc, generator=g).expand(B, L, Cc).contiguous()
    umask = torch.zeros(B, L, dtype=torch.int64)
    umask[:, 0] = 1
    return tuple(t.to(device) for t in (noise, ctx, mask, uctx, umask))
