"""ORACLE (test infrastructure, never on the product path): fp32 PyTorch restatement of the
PixArt-alpha DiT that the reference instantiates as `diffusers.Transformer2DModel`
(reference main.py:46, 168-176, 367-369) and calls at models.py:38-43, predictor.py:33-39,
main.py:481-494.

The arithmetic lives in the un-vendored third-party package `diffusers` (only pin in the reference:
check_min_version("0.26.0.dev0"), main.py:68), which is absent from /root/reference and not
installable here, so this file restates its published algorithm for
`norm_type="ada_norm_single"` with diffusers' parameter names, so that a real
PixArt checkpoint `state_dict` would load.  PARITY UNPINNED against diffusers itself: the reference
ships no tests or golden vectors for the transformer; building blocks are cross-checked against
independent implementations in tests/test_oracle_dit.py (F.layer_norm, F.scaled_dot_product_attention,
F.gelu, transformers' ViT-MAE sincos embedding).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class DiTConfig:
    num_attention_heads: int = 16
    attention_head_dim: int = 72
    num_layers: int = 28
    in_channels: int = 4
    out_channels: int = 8
    patch_size: int = 2
    sample_size: int = 64
    caption_channels: int = 4096
    norm_eps: float = 1e-6
    # PixArt-XL-2-1024 would use sample_size=128 (interpolation_scale 2); micro-conditioning stays off
    # because the reference always passes resolution/aspect_ratio = None (predictor.py:28).

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim

    @property
    def interpolation_scale(self) -> int:
        return max(self.sample_size // 64, 1)

    def to_dict(self):
        return asdict(self)


PIXART_XL_2_512 = DiTConfig()
PIXART_XL_2_1024 = DiTConfig(sample_size=128)
TINY = DiTConfig(num_attention_heads=4, attention_head_dim=64, num_layers=2, sample_size=32, caption_channels=4096)


# ---- positional / timestep embeddings (Appendix A, steps 2-3) -----------------------------------
def sincos_1d(embed_dim: int, pos: np.ndarray) -> np.ndarray:
    omega = np.arange(embed_dim // 2, dtype=np.float64) / (embed_dim / 2.0)
    omega = 1.0 / 10000 ** omega
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def sincos_2d(embed_dim: int, grid_size: int, base_size: int, interpolation_scale: float) -> np.ndarray:
    gh = np.arange(grid_size, dtype=np.float32) / (grid_size / base_size) / interpolation_scale
    gw = np.arange(grid_size, dtype=np.float32) / (grid_size / base_size) / interpolation_scale
    grid = np.stack(np.meshgrid(gw, gh), axis=0).reshape(2, 1, grid_size, grid_size)  # w first
    emb_h = sincos_1d(embed_dim // 2, grid[0])
    emb_w = sincos_1d(embed_dim // 2, grid[1])
    return np.concatenate([emb_h, emb_w], axis=1)  # [grid*grid, D]


def timestep_sinusoid(t: torch.Tensor, dim: int = 256) -> torch.Tensor:
    """Timesteps(256, flip_sin_to_cos=True, downscale_freq_shift=0): cat[cos, sin]."""
    half = dim // 2
    freq = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    arg = t.float()[:, None] * freq[None]
    return torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1)


class _TimestepEmbedder(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.linear_1 = nn.Linear(256, d)
        self.linear_2 = nn.Linear(d, d)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class _Emb(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.timestep_embedder = _TimestepEmbedder(d)


class AdaLayerNormSingle(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.emb = _Emb(d)
        self.linear = nn.Linear(d, 6 * d)

    def forward(self, timestep, dtype):
        embedded = self.emb.timestep_embedder(timestep_sinusoid(timestep).to(dtype))
        return self.linear(F.silu(embedded)), embedded


class CaptionProjection(nn.Module):
    def __init__(self, c_in, d):
        super().__init__()
        self.linear_1 = nn.Linear(c_in, d)
        self.linear_2 = nn.Linear(d, d)

    def forward(self, x):
        return self.linear_2(F.gelu(self.linear_1(x), approximate="tanh"))


class PatchEmbed(nn.Module):
    def __init__(self, cfg: DiTConfig):
        super().__init__()
        d, p = cfg.inner_dim, cfg.patch_size
        self.proj = nn.Conv2d(cfg.in_channels, d, kernel_size=p, stride=p, bias=True)
        self.cfg = cfg
        grid = cfg.sample_size // p
        pe = sincos_2d(d, grid, base_size=grid, interpolation_scale=cfg.interpolation_scale)
        self.register_buffer("pos_embed", torch.from_numpy(pe).float().unsqueeze(0), persistent=False)

    def forward(self, latent):
        x = self.proj(latent).flatten(2).transpose(1, 2)
        gh = latent.shape[-2] // self.cfg.patch_size
        if gh * gh != self.pos_embed.shape[1]:
            base = self.cfg.sample_size // self.cfg.patch_size
            pe = sincos_2d(x.shape[-1], gh, base_size=base, interpolation_scale=self.cfg.interpolation_scale)
            pe = torch.from_numpy(pe).float().unsqueeze(0).to(x.device)
        else:
            pe = self.pos_embed
        return (x + pe).to(x.dtype)


class Attention(nn.Module):
    # bench.py's baselines may switch the softmax(QK^T)V evaluation to torch's fused SDPA (same function, no S
    # materialised: the stand-in for the reference's xformers call, main.py:177-178); parity tests keep the explicit form
    use_sdpa = False

    def __init__(self, d, heads):
        super().__init__()
        self.heads = heads
        self.to_q = nn.Linear(d, d)
        self.to_k = nn.Linear(d, d)
        self.to_v = nn.Linear(d, d)
        self.to_out = nn.ModuleList([nn.Linear(d, d), nn.Dropout(0.0)])

    def forward(self, x, ctx=None, bias=None):
        ctx = x if ctx is None else ctx
        B, N, D = x.shape
        h = self.heads
        q = self.to_q(x).view(B, N, h, D // h).transpose(1, 2)
        k = self.to_k(ctx).view(B, -1, h, D // h).transpose(1, 2)
        v = self.to_v(ctx).view(B, -1, h, D // h).transpose(1, 2)
        if self.use_sdpa:
            m = None if bias is None else bias[:, None].to(q.dtype).expand(B, h, N, k.shape[2])
            o = F.scaled_dot_product_attention(q, k, v, attn_mask=m, scale=(D // h) ** -0.5)
        else:
            s = torch.matmul(q, k.transpose(-1, -2)) * (D // h) ** -0.5
            if bias is not None:
                s = s + bias[:, None]  # [B,1,1,L] broadcast over heads and queries
            o = torch.matmul(torch.softmax(s, dim=-1), v)
        o = o.transpose(1, 2).reshape(B, N, D)
        return self.to_out[0](o)


class _GELUProj(nn.Module):
    def __init__(self, d, d_out):
        super().__init__()
        self.proj = nn.Linear(d, d_out)

    def forward(self, x):
        return F.gelu(self.proj(x), approximate="tanh")


class FeedForward(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.net = nn.ModuleList([_GELUProj(d, 4 * d), nn.Dropout(0.0), nn.Linear(4 * d, d)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class BasicTransformerBlock(nn.Module):
    def __init__(self, cfg: DiTConfig):
        super().__init__()
        d = cfg.inner_dim
        self.eps = cfg.norm_eps
        self.scale_shift_table = nn.Parameter(torch.randn(6, d) / d ** 0.5)
        self.attn1 = Attention(d, cfg.num_attention_heads)
        self.attn2 = Attention(d, cfg.num_attention_heads)
        self.ff = FeedForward(d)

    def forward(self, x, ctx, bias, timestep_mod):
        B, _, D = x.shape
        mod = self.scale_shift_table[None] + timestep_mod.reshape(B, 6, -1)
        shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = mod.chunk(6, dim=1)
        h = F.layer_norm(x, (D,), eps=self.eps) * (1 + scale_msa) + shift_msa
        x = x + gate_msa * self.attn1(h)
        x = x + self.attn2(x, ctx, bias)
        h = F.layer_norm(x, (D,), eps=self.eps) * (1 + scale_mlp) + shift_mlp
        x = x + gate_mlp * self.ff(h)
        return x


class OracleDiT(nn.Module):
    """Call surface == what the reference touches on `unet` / `unet_fake` / `unet_sd`
    (SURVEY.md section 8b): __call__(sample, timestep=, added_cond_kwargs=, encoder_hidden_states=,
    encoder_attention_mask=, return_dict=False) -> (Tensor[B, 2*C_in, H, W],)."""

    def __init__(self, cfg: DiTConfig = PIXART_XL_2_512):
        super().__init__()
        self.cfg = cfg
        d = cfg.inner_dim
        self.pos_embed = PatchEmbed(cfg)
        self.adaln_single = AdaLayerNormSingle(d)
        self.caption_projection = CaptionProjection(cfg.caption_channels, d)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(cfg) for _ in range(cfg.num_layers)])
        self.scale_shift_table = nn.Parameter(torch.randn(2, d) / d ** 0.5)
        self.proj_out = nn.Linear(d, cfg.patch_size * cfg.patch_size * cfg.out_channels)
        self.capture = None  # optional dict: per-block outputs for parity tests

    def forward(self, hidden_states, timestep=None, added_cond_kwargs=None, encoder_hidden_states=None,
                encoder_attention_mask=None, return_dict=False):
        cfg = self.cfg
        wdtype = self.proj_out.weight.dtype
        # D10: the reference feeds fp16 latents to fp32 weights and relies on autocast;
        # the fp32 oracle upcasts at model entry.
        x_in = hidden_states.to(wdtype)
        ctx_in = encoder_hidden_states.to(wdtype)
        bias = None
        if encoder_attention_mask is not None and encoder_attention_mask.ndim == 2:
            bias = ((1 - encoder_attention_mask.to(wdtype)) * -10000.0).unsqueeze(1)  # [B,1,L]
        B, _, H, W = x_in.shape
        gh, gw = H // cfg.patch_size, W // cfg.patch_size
        x = self.pos_embed(x_in)
        timestep_mod, embedded = self.adaln_single(timestep, wdtype)
        ctx = self.caption_projection(ctx_in).view(B, -1, x.shape[-1])
        for i, blk in enumerate(self.transformer_blocks):
            x = blk(x, ctx, bias, timestep_mod)
            if self.capture is not None:
                self.capture[f"block{i}"] = x.detach().clone()
        shift, scale = (self.scale_shift_table[None] + embedded[:, None]).chunk(2, dim=1)
        x = F.layer_norm(x, (x.shape[-1],), eps=cfg.norm_eps) * (1 + scale) + shift
        x = self.proj_out(x)
        p, c = cfg.patch_size, cfg.out_channels
        x = x.reshape(B, gh, gw, p, p, c)
        x = torch.einsum("nhwpqc->nchpwq", x).reshape(B, c, gh * p, gw * p)
        return (x,)

    # no-op members the reference calls on its models (main.py:177-178, 191-193)
    def enable_xformers_memory_efficient_attention(self):
        return None

    def enable_gradient_checkpointing(self):
        return None


def count_params(cfg: DiTConfig) -> int:
    d = cfg.inner_dim
    per_block = 16 * d * d + 19 * d
    p = cfg.patch_size
    return (cfg.num_layers * per_block + (cfg.in_channels * p * p * d + d) + (256 * d + d + d * d + d)
            + (d * 6 * d + 6 * d) + (cfg.caption_channels * d + d + d * d + d) + (d * p * p * cfg.out_channels
            + p * p * cfg.out_channels) + 2 * d)
