"""ORACLE (test infrastructure): LoRA with peft's semantics on the oracle DiT's nn.Linear layers.

The reference imports peft (main.py:43, utils.py:18) but never attaches an adapter;
north_star asks for a LoRA student, so the oracle restates peft's published forward
    y = base(x) + (lora_alpha / r) * lora_B(lora_A(x)),   A ~ kaiming_uniform(a=sqrt 5), B = 0
with peft's parameter names (`<linear>.lora_A.default.weight`, `<linear>.lora_B.default.weight`), the
names utils.py:76-93 converts for the kohya export.  peft itself is not installable here: parity unpinned
against the package, pinned by construction against the formula above.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn


class _Adapter(nn.Module):
    def __init__(self, fin, fout):
        super().__init__()
        self.default = nn.Linear(fin, fout, bias=False)


class LoRALinear(nn.Module):
    def __init__(self, base: nn.Linear, r: int, alpha: float):
        super().__init__()
        self.weight, self.bias = base.weight, base.bias
        self.lora_A = _Adapter(base.in_features, r)
        self.lora_B = _Adapter(r, base.out_features)
        nn.init.kaiming_uniform_(self.lora_A.default.weight, a=math.sqrt(5))
        nn.init.zeros_(self.lora_B.default.weight)
        self.scaling = alpha / r

    def forward(self, x):
        y = torch.nn.functional.linear(x, self.weight, self.bias)
        return y + self.scaling * self.lora_B.default(self.lora_A.default(x))


def add_lora(model, r=64, alpha=None, targets=("attn1.to_q", "attn1.to_k", "attn1.to_v", "attn1.to_out.0",
                                               "attn2.to_q", "attn2.to_k", "attn2.to_v", "attn2.to_out.0",
                                               "ff.net.0.proj", "ff.net.2"), freeze_base=True):
    alpha = r if alpha is None else alpha
    if freeze_base:
        for p in model.parameters():
            p.requires_grad_(False)
    for blk in model.transformer_blocks:
        for t in targets:
            parts = t.split(".")
            parent = blk
            for q in parts[:-1]:
                parent = parent[int(q)] if q.isdigit() else getattr(parent, q)
            last = parts[-1]
            base = parent[int(last)] if last.isdigit() else getattr(parent, last)
            wrapped = LoRALinear(base, r, alpha)
            if last.isdigit():
                parent[int(last)] = wrapped
            else:
                setattr(parent, last, wrapped)
    return model
