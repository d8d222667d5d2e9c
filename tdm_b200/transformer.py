"""Drop-in for the `diffusers.Transformer2DModel` (PixArt-alpha DiT, norm_type="ada_norm_single") that
the reference builds three times (student main.py:168-170, fake score 174-176, teacher 367-369) and calls
at models.py:38-43, predictor.py:33-39, main.py:481-494.  Same call surface, diffusers
parameter names, but every FLOP runs in libtdm_b200's sm_100a kernels.

Memory layout (HBM): all parameters of a model live in ONE flat fp32 arena (each nn.Parameter is a view),
mirrored by one flat bf16 arena the tensor-core kernels read and one flat fp32 gradient arena the
backward kernels accumulate into.  q/k/v (and cross k/v) weights are adjacent, so the fused QKV GEMM
reads one [3D, D] matrix without any copy; data-parallel all-reduce and the optimizer work on the flat
arenas directly.  LoRA adapters (peft semantics) have their own small arenas.
"""
from __future__ import annotations

import ctypes as C
import math
from collections import OrderedDict

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import check, ptr, stream_handle

ALIGN = 128  # elements: 512 B in fp32, 256 B in bf16 (TMA needs 16 B, the block workspace 256 B)

DEFAULT_CONFIG = dict(num_attention_heads=16, attention_head_dim=72, num_layers=28, in_channels=4, out_channels=8,
                      patch_size=2, sample_size=64, caption_channels=4096, norm_eps=1e-6)
TINY_CONFIG = dict(DEFAULT_CONFIG, num_attention_heads=4, attention_head_dim=64, num_layers=2, sample_size=32)

LORA_TARGETS_DEFAULT = ("attn1.to_q", "attn1.to_k", "attn1.to_v", "attn1.to_out.0", "attn2.to_q", "attn2.to_k",
                        "attn2.to_v", "attn2.to_out.0", "ff.net.0.proj", "ff.net.2")


class _Arena:
    """name -> (offset, shape) layout over flat fp32 / bf16 / grad buffers."""

    def __init__(self):
        self.layout = OrderedDict()
        self.size = 0

    def add(self, name, shape):
        n = int(math.prod(shape))
        self.layout[name] = (self.size, tuple(shape))
        self.size += (n + ALIGN - 1) // ALIGN * ALIGN

    def view(self, flat, name):
        off, shape = self.layout[name]
        return flat[off:off + int(math.prod(shape))].view(shape)


def _sincos_pos_embed(D, grid, base, scale, device):
    """get_2d_sincos_pos_embed, fp64 math like diffusers' numpy code."""
    def one_d(dim, pos):
        omega = 1.0 / 10000 ** (torch.arange(dim // 2, dtype=torch.float64) / (dim / 2.0))
        out = pos.reshape(-1).double()[:, None] * omega[None]
        return torch.cat([out.sin(), out.cos()], dim=1)
    g = torch.arange(grid, dtype=torch.float32) / (grid / base) / scale
    gw, gh = torch.meshgrid(g, g, indexing="xy")  # np.meshgrid(grid_w, grid_h): w first
    return torch.cat([one_d(D // 2, gw), one_d(D // 2, gh)], dim=1).float().to(device)


class _Saved:
    """What one training forward keeps for its backward."""
    __slots__ = ("B", "N", "L", "HW", "shape", "patches", "tsin", "t_pre1", "t_h", "emb_pre", "semb", "tmod", "ctx_in",
                 "c_pre", "c1", "ctx", "ctx_bias", "xs", "ws", "modf", "hf", "meanf", "rstdf")


class Transformer2DModel(nn.Module):
    def __init__(self, config=None, device="cuda", seed=None, init_weights=True, **overrides):
        super().__init__()
        cfg = dict(DEFAULT_CONFIG)
        cfg.update(config or {})
        cfg.update(overrides)
        self.config_dict = cfg
        self.H = cfg["num_attention_heads"]
        self.dh = cfg["attention_head_dim"]
        self.D = self.H * self.dh
        self.depth = cfg["num_layers"]
        self.Cin, self.Cout, self.p = cfg["in_channels"], cfg["out_channels"], cfg["patch_size"]
        self.sample_size = cfg["sample_size"]
        self.Cc = cfg["caption_channels"]
        self.eps = cfg["norm_eps"]
        self.Dff = 4 * self.D
        self.lora_r = 0
        self.lora_scale = 0.0
        self.gradient_checkpointing = False
        D, Dff, p = self.D, self.Dff, self.p

        A = _Arena()
        A.add("pos_embed.proj.weight", (D, self.Cin, p, p))
        A.add("pos_embed.proj.bias", (D,))
        A.add("adaln_single.emb.timestep_embedder.linear_1.weight", (D, 256))
        A.add("adaln_single.emb.timestep_embedder.linear_1.bias", (D,))
        A.add("adaln_single.emb.timestep_embedder.linear_2.weight", (D, D))
        A.add("adaln_single.emb.timestep_embedder.linear_2.bias", (D,))
        A.add("adaln_single.linear.weight", (6 * D, D))
        A.add("adaln_single.linear.bias", (6 * D,))
        A.add("caption_projection.linear_1.weight", (D, self.Cc))
        A.add("caption_projection.linear_1.bias", (D,))
        A.add("caption_projection.linear_2.weight", (D, D))
        A.add("caption_projection.linear_2.bias", (D,))
        for i in range(self.depth):
            b = f"transformer_blocks.{i}."
            A.add(b + "scale_shift_table", (6, D))
            # q,k,v adjacent and un-padded between them (D*D is a multiple of ALIGN) -> one [3D, D] matrix
            for n in ("attn1.to_q", "attn1.to_k", "attn1.to_v"):
                A.add(b + n + ".weight", (D, D))
            for n in ("attn1.to_q", "attn1.to_k", "attn1.to_v"):
                A.add(b + n + ".bias", (D,))
            A.add(b + "attn1.to_out.0.weight", (D, D))
            A.add(b + "attn1.to_out.0.bias", (D,))
            A.add(b + "attn2.to_q.weight", (D, D))
            A.add(b + "attn2.to_q.bias", (D,))
            for n in ("attn2.to_k", "attn2.to_v"):
                A.add(b + n + ".weight", (D, D))
            for n in ("attn2.to_k", "attn2.to_v"):
                A.add(b + n + ".bias", (D,))
            A.add(b + "attn2.to_out.0.weight", (D, D))
            A.add(b + "attn2.to_out.0.bias", (D,))
            A.add(b + "ff.net.0.proj.weight", (Dff, D))
            A.add(b + "ff.net.0.proj.bias", (Dff,))
            A.add(b + "ff.net.2.weight", (D, Dff))
            A.add(b + "ff.net.2.bias", (D,))
        A.add("scale_shift_table", (2, D))
        A.add("proj_out.weight", (p * p * self.Cout, D))
        A.add("proj_out.bias", (p * p * self.Cout,))
        assert D % ALIGN == 0, "inner dim must be a multiple of 128"
        self._arena = A
        self._lora_arena = None
        self._flat = torch.zeros(A.size, device=device, dtype=torch.float32)
        self._flat16 = None
        self._flat16_version = -1
        self._grad = None
        self._lflat = self._lflat16 = self._lgrad = None
        self._lflat16_version = -1
        self._bind_params()
        if init_weights:
            self._init_weights(seed)
        grid = self.sample_size // p
        self._pos_embed = _sincos_pos_embed(D, grid, grid, max(self.sample_size // 64, 1), device)
        self._pos_cache = {grid: self._pos_embed}
        self._anchor = torch.zeros((), device=device, requires_grad=True)  # ties custom backward into autograd
        self._block_structs = None
        self._scratch = None
        self._infer_ws = None
        # no-grad forwards of a LoRA model read W_eff = W + (alpha/r) B A, merged once per optimizer step
        self.lora_merge_infer = True
        self._flat16_eff = None
        self._eff_key = None
        self._merged_structs = None
        self._merge_table = None
        self._update_count = 0

    # ------------------------------------------------------------------ parameters
    def _set_param(self, name, tensor, requires_grad=True):
        parts = name.split(".")
        mod = self
        for part in parts[:-1]:
            if part not in mod._modules:
                mod.add_module(part, nn.Module())
            mod = mod._modules[part]
        mod.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=requires_grad))

    def _bind_params(self):
        old = {n: p.requires_grad for n, p in self.named_parameters()}
        for name in self._arena.layout:
            self._set_param(name, self._arena.view(self._flat, name), old.get(name, True))
        if self._lora_arena is not None:
            for name in self._lora_arena.layout:
                self._set_param(name, self._lora_arena.view(self._lflat, name), old.get(name, True))
        self._block_structs = None

    def _init_weights(self, seed):
        """torch default inits of the oracle's modules (nn.Linear / nn.Conv2d kaiming-uniform(a=sqrt 5),
        tables randn/sqrt(D))."""
        gen = torch.Generator(device="cpu")
        if seed is not None:
            gen.manual_seed(seed)
        else:
            gen.manual_seed(torch.initial_seed() % (2 ** 63))
        with torch.no_grad():
            for name, (off, shape) in self._arena.layout.items():
                p = self._arena.view(self._flat, name)
                if name.endswith("scale_shift_table"):
                    p.copy_(torch.randn(shape, generator=gen) / shape[1] ** 0.5)
                elif name.endswith(".weight"):
                    fan_in = int(math.prod(shape[1:]))
                    bound = 1.0 / math.sqrt(fan_in)
                    p.copy_((torch.rand(shape, generator=gen) * 2 - 1) * bound)
                else:  # bias: uniform(-1/sqrt(fan_in), 1/sqrt(fan_in)) of its weight
                    woff, wshape = self._arena.layout[name[:-4] + "weight"]
                    bound = 1.0 / math.sqrt(int(math.prod(wshape[1:])))
                    p.copy_((torch.rand(shape, generator=gen) * 2 - 1) * bound)

    def _apply(self, fn, recurse=True):
        new = fn(self._flat)
        if new.dtype != torch.float32:
            raise _lib.TdmError("Transformer2DModel keeps fp32 master parameters; compute runs in bf16 internally")
        if new is not self._flat:
            if not new.is_cuda:
                raise _lib.TdmError("Transformer2DModel lives on a CUDA device (no CPU fallback)")
            self._flat = new
            if self._lflat is not None:
                self._lflat = fn(self._lflat)
            self._pos_embed = fn(self._pos_embed)
            self._pos_cache = {}
            self._anchor = torch.zeros((), device=new.device, requires_grad=True)
            self._flat16 = self._grad = self._lflat16 = self._lgrad = None
            self._flat16_eff = self._eff_key = self._merged_structs = self._merge_table = None
            self._flat16_version = self._lflat16_version = -1
            self._scratch = self._infer_ws = None
            self._bind_params()
        return self

    @property
    def device(self):
        return self._flat.device

    @property
    def dtype(self):
        return torch.float32

    @classmethod
    def from_config(cls, config, **kw):
        return cls(config, **kw)

    _pretrained_provider = None

    @classmethod
    def from_pretrained(cls, path, subfolder=None, revision=None, variant=None, **kw):
        """main.py:168-176, 367-369.  No network here: a provider (callable -> model) must be registered by the
        harness (random-init weights per north_star), otherwise this is an error, never a silent download."""
        if cls._pretrained_provider is None:
            raise _lib.TdmError(f"from_pretrained({path!r}): no weight provider registered (offline build)")
        return cls._pretrained_provider(path, subfolder=subfolder, revision=revision, variant=variant, **kw)

    def enable_xformers_memory_efficient_attention(self):  # main.py:177-178, 372
        return None  # the tcgen05 attention kernel never materialises S either

    def enable_gradient_checkpointing(self):  # main.py:191-193
        self.gradient_checkpointing = True  # accepted; 180 GB HBM holds the activations

    def load_oracle_state(self, state_dict):
        """Copy parameters by diffusers name (e.g. from the oracle DiT or a real checkpoint)."""
        with torch.no_grad():
            own = dict(self.named_parameters())
            for k, v in state_dict.items():
                if k in own:
                    own[k].copy_(v.to(own[k].device, torch.float32))
        return self

    def copy_weights_from(self, other: "Transformer2DModel"):
        with torch.no_grad():
            self._flat.copy_(other._flat)
        return self

    # ------------------------------------------------------------------ LoRA (peft semantics)
    def add_lora(self, rank=64, alpha=None, targets=LORA_TARGETS_DEFAULT, seed=None, freeze_base=True):
        """y = W x + b + (alpha/r) * B(A x); A ~ kaiming_uniform(a=sqrt 5), B = 0; names follow peft
        (`...to_q.lora_A.default.weight`), which is what utils.py:76-93 exports."""
        if rank != 64:
            raise _lib.TdmError("the fused LoRA epilogue supports rank 64")
        D, Dff = self.D, self.Dff
        dims = {"attn1.to_q": (D, D), "attn1.to_k": (D, D), "attn1.to_v": (D, D), "attn1.to_out.0": (D, D),
                "attn2.to_q": (D, D), "attn2.to_k": (D, D), "attn2.to_v": (D, D), "attn2.to_out.0": (D, D),
                "ff.net.0.proj": (Dff, D), "ff.net.2": (D, Dff)}
        groups = [("attn1.to_q", "attn1.to_k", "attn1.to_v"), ("attn2.to_k", "attn2.to_v")]
        for grp in groups:
            present = [t in targets for t in grp]
            if any(present) and not all(present):
                raise _lib.TdmError(f"LoRA targets must include all of {grp} or none (fused projection)")
        L = _Arena()
        for i in range(self.depth):
            b = f"transformer_blocks.{i}."
            for t in dims:  # A matrices of a fused group adjacent, then their B matrices adjacent
                if t in targets:
                    L.add(b + t + ".lora_A.default.weight", (rank, dims[t][1]))
            for t in dims:
                if t in targets:
                    L.add(b + t + ".lora_B.default.weight", (dims[t][0], rank))
        self._lora_arena = L
        self.lora_targets = tuple(targets)
        self.lora_r = rank
        self.lora_scale = float(alpha if alpha is not None else rank) / rank
        import types
        self.peft_config = {"default": types.SimpleNamespace(r=rank, lora_alpha=float(alpha if alpha is not None else rank),
                                                              target_modules=list(targets))}
        self._lflat = torch.zeros(L.size, device=self.device, dtype=torch.float32)
        gen = torch.Generator(device="cpu")
        gen.manual_seed(seed if seed is not None else torch.initial_seed() % (2 ** 63))
        with torch.no_grad():
            for name, (off, shape) in L.layout.items():
                if ".lora_A." in name:
                    bound = 1.0 / math.sqrt(shape[1])  # kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), +)
                    L.view(self._lflat, name).copy_((torch.rand(shape, generator=gen) * 2 - 1) * bound)
        if freeze_base:
            for n, p in self.named_parameters():
                p.requires_grad_(False)
        self._bind_params()
        for n, p in self.named_parameters():
            if ".lora_" in n:
                p.requires_grad_(True)
        self._lflat16 = self._lgrad = None
        self._flat16_eff = self._eff_key = self._merged_structs = self._merge_table = None
        return self

    def peft_state_dict(self, adapter_name="default"):
        """What peft.get_peft_model_state_dict returns for a wrapped model: adapter weights only, keys prefixed with
        `base_model.model.` and with the adapter name dropped (`...to_q.lora_A.weight`)."""
        tag = f".{adapter_name}.weight"
        return {"base_model.model." + n[:-len(tag)] + ".weight": p.detach()
                for n, p in self.named_parameters() if ".lora_" in n and n.endswith(tag)}

    def lora_parameters(self):
        return [p for n, p in self.named_parameters() if ".lora_" in n]

    # ------------------------------------------------------------------ arenas
    @property
    def train_base(self) -> bool:
        return any(p.requires_grad for n, p in self.named_parameters() if ".lora_" not in n)

    def _sync_bf16(self):
        if self._flat16 is None:
            self._flat16 = torch.empty(self._arena.size, device=self.device, dtype=torch.bfloat16)
            self._block_structs = None
        if self._flat._version != self._flat16_version:
            check(_lib.lib().tdm_cast_f32_bf16(ptr(self._flat), ptr(self._flat16), self._arena.size, stream_handle()))
            self._flat16_version = self._flat._version
        if self._lflat is not None:
            if self._lflat16 is None:
                self._lflat16 = torch.empty(self._lora_arena.size, device=self.device, dtype=torch.bfloat16)
                self._block_structs = None
            if self._lflat._version != self._lflat16_version:
                check(_lib.lib().tdm_cast_f32_bf16(ptr(self._lflat), ptr(self._lflat16), self._lora_arena.size,
                                                   stream_handle()))
                self._lflat16_version = self._lflat._version

    def _ensure_grads(self):
        """Flat gradient arenas; each trainable parameter's .grad is a view (re-attached and zeroed after
        optimizer.zero_grad(set_to_none=True))."""
        tb = self.train_base
        if tb and self._grad is None:
            self._grad = torch.zeros(self._arena.size, device=self.device, dtype=torch.float32)
            self._block_structs = None
        if self._lflat is not None and self._lgrad is None:
            self._lgrad = torch.zeros(self._lora_arena.size, device=self.device, dtype=torch.float32)
            self._block_structs = None
        detached = False
        for n, p in self.named_parameters():
            if p.requires_grad and p.grad is None:
                detached = True
                break
        if detached:
            if tb:
                self._grad.zero_()
            if self._lgrad is not None:
                self._lgrad.zero_()
            for n, p in self.named_parameters():
                if not p.requires_grad:
                    continue
                if ".lora_" in n:
                    p.grad = self._lora_arena.view(self._lgrad, n)
                else:
                    p.grad = self._arena.view(self._grad, n)

    def flat_grads(self):
        """The flat gradient buffers (for bucketed all-reduce / fused optimizers)."""
        out = []
        if self.train_base and self._grad is not None:
            out.append(self._grad)
        if self._lgrad is not None:
            out.append(self._lgrad)
        return out

    def flat_params(self):
        out = []
        if self.train_base:
            out.append(self._flat)
        if self._lflat is not None:
            out.append(self._lflat)
        return out

    def _w16(self, name):
        return self._arena.view(self._flat16, name)

    def _w32(self, name):
        return self._arena.view(self._flat, name)

    def _g32(self, name):
        return self._arena.view(self._grad, name)

    # ------------------------------------------------------------------ per-block C structs
    def _build_block_structs(self):
        self._sync_bf16()
        ws, gs = [], []
        tb = self.train_base and self._grad is not None
        for i in range(self.depth):
            b = f"transformer_blocks.{i}."
            w = _lib.TdmBlockWeights()
            g = _lib.TdmBlockGrads()

            def base(field, wname, bname=None):
                setattr(w, "w_" + field, ptr(self._w16(b + wname + ".weight")))
                setattr(w, "b_" + field, ptr(self._w32(b + (bname or wname) + ".bias")))
                if tb:
                    setattr(g, "w_" + field, ptr(self._g32(b + wname + ".weight")))
                    setattr(g, "b_" + field, ptr(self._g32(b + (bname or wname) + ".bias")))

            w.table = ptr(self._w32(b + "scale_shift_table"))
            if tb:
                g.table = ptr(self._g32(b + "scale_shift_table"))
            base("qkv", "attn1.to_q")
            base("o", "attn1.to_out.0")
            base("q2", "attn2.to_q")
            base("kv2", "attn2.to_k")
            base("o2", "attn2.to_out.0")
            base("ff1", "ff.net.0.proj")
            base("ff2", "ff.net.2")
            if self._lflat is not None:
                LA = self._lora_arena
                for field, first in (("qkv", "attn1.to_q"), ("o", "attn1.to_out.0"), ("q2", "attn2.to_q"),
                                     ("kv2", "attn2.to_k"), ("o2", "attn2.to_out.0"), ("ff1", "ff.net.0.proj"),
                                     ("ff2", "ff.net.2")):
                    an, bn = b + first + ".lora_A.default.weight", b + first + ".lora_B.default.weight"
                    if an in LA.layout:
                        setattr(w, "lora_a_" + field, ptr(LA.view(self._lflat16, an)))
                        setattr(w, "lora_b_" + field, ptr(LA.view(self._lflat16, bn)))
                        if self._lgrad is not None:
                            setattr(g, "lora_a_" + field, ptr(LA.view(self._lgrad, an)))
                            setattr(g, "lora_b_" + field, ptr(LA.view(self._lgrad, bn)))
            ws.append(w)
            gs.append(g)
        self._block_structs = (ws, gs)

    def mark_updated(self):
        """Called by optimizers that write the arenas through raw pointers (torch's version counters do not see
        those writes): invalidates the merged LoRA weights."""
        self._update_count += 1

    def _refresh_merged(self):
        """W_eff16 = bf16(W32 + (alpha/r) * B A) for every adapted linear in ONE launch (csrc/lora_merge.cu).  Redone only
        after the adapters changed."""
        key = (self._lflat._version, self._flat._version, self._update_count)
        if self._flat16_eff is not None and self._eff_key == key:
            return
        self._sync_bf16()
        if self._flat16_eff is None:
            self._flat16_eff = torch.empty_like(self._flat16)
            self._merged_structs = self._merge_table = None
        if getattr(self, "_eff_base_version", None) != self._flat._version:
            # everything that is NOT an adapted weight (biases, tables, embedders) is a plain copy of the bf16 arena; with a
            # frozen base (the LoRA student) that copy is made once, not 1.2 GB per optimiser step
            self._flat16_eff.copy_(self._flat16)
            self._eff_base_version = self._flat._version
        if self._merge_table is None:
            self._build_merge_table()
        table, n_prob, n_tiles = self._merge_table
        check(_lib.lib().tdm_lora_merge(ptr(table), n_prob, n_tiles, float(self.lora_scale), stream_handle()))
        self._eff_key = key
        if self._merged_structs is None:
            ws = []
            for i in range(self.depth):
                b = f"transformer_blocks.{i}."
                w = _lib.TdmBlockWeights()
                w.table = ptr(self._w32(b + "scale_shift_table"))
                for field, wname in (("qkv", "attn1.to_q"), ("o", "attn1.to_out.0"), ("q2", "attn2.to_q"),
                                     ("kv2", "attn2.to_k"), ("o2", "attn2.to_out.0"), ("ff1", "ff.net.0.proj"),
                                     ("ff2", "ff.net.2")):
                    setattr(w, "w_" + field, ptr(self._arena.view(self._flat16_eff, b + wname + ".weight")))
                    setattr(w, "b_" + field, ptr(self._w32(b + wname + ".bias")))
                ws.append(w)
            self._merged_structs = ws

    def _build_merge_table(self):
        """Device-side table of (fp32 master, merged bf16 copy, LoRA A, LoRA B, shape) for every adapted linear: the arenas
        never move, so it is built once (TdmMergeProblem, include/tdm_b200.h)."""
        import struct
        LA = self._lora_arena
        rows, tile = [], 0
        for i in range(self.depth):
            b = f"transformer_blocks.{i}."
            for t in self.lora_targets:
                an, bn = b + t + ".lora_A.default.weight", b + t + ".lora_B.default.weight"
                w32 = self._arena.view(self._flat, b + t + ".weight")
                weff = self._arena.view(self._flat16_eff, b + t + ".weight")
                n_out, k_in = w32.shape
                if n_out % 128 or k_in % 128:
                    raise _lib.TdmError("LoRA merge needs weight dimensions that are multiples of 128")
                rows.append(struct.pack("<QQQQiiii", w32.data_ptr(), weff.data_ptr(), LA.view(self._lflat16, an).data_ptr(),
                                        LA.view(self._lflat16, bn).data_ptr(), n_out, k_in, tile, 0))
                tile += (n_out // 128) * (k_in // 128)
        raw = torch.frombuffer(bytearray(b"".join(rows)), dtype=torch.uint8).to(self.device)
        self._merge_table = (raw, len(rows), tile)

    def _shape(self, B, N, L, lora=True):
        s = _lib.TdmDitShape()
        s.B, s.N, s.D, s.H, s.L, s.Dff = B, N, self.D, self.H, L, self.Dff
        s.lora_r = self.lora_r if (self._lflat is not None and lora) else 0
        s.lora_scale = self.lora_scale
        s.eps = self.eps
        s.train_base = int(self.train_base)
        return s

    # ------------------------------------------------------------------ forward
    def _pos(self, gh):
        if gh not in self._pos_cache:
            base = self.sample_size // self.p
            self._pos_cache[gh] = _sincos_pos_embed(self.D, gh, base, max(self.sample_size // 64, 1), self.device)
        return self._pos_cache[gh]

    def _forward_impl(self, latent, timestep, ctx_in, mask, save):
        lib = _lib.lib()
        st = stream_handle()
        dev = self.device
        D, p = self.D, self.p
        if not latent.is_cuda:
            raise _lib.TdmError("Transformer2DModel: inputs must be CUDA tensors (no CPU fallback)")
        B, _, Hh, Ww = latent.shape
        gh, gw = Hh // p, Ww // p
        N = gh * gw
        M = B * N
        L = ctx_in.shape[1]
        self._sync_bf16()
        if save:
            self._ensure_grads()
        if self._block_structs is None:
            self._build_block_structs()
        bw, _ = self._block_structs
        merged = (not save) and self._lflat is not None and self.lora_merge_infer
        if merged:
            self._refresh_merged()
            bw = self._merged_structs
        f32 = dict(device=dev, dtype=torch.float32)
        b16 = dict(device=dev, dtype=torch.bfloat16)

        # K12 patch embed + pos embed: conv(k=2,s=2) as a K=16 contraction, pos_embed as the residual operand
        patches = ops.patchify(latent, p)
        x = torch.empty(M, D, **f32)
        pos = self._pos(gh) if gh == gw else None
        if pos is None:
            raise _lib.TdmError("non-square latents are not supported")
        ops.gemm(patches, self._w16("pos_embed.proj.weight").view(D, -1), bias=self._w32("pos_embed.proj.bias"),
                 res=pos, res_row_mod=N, out_f32=x)
        # K13 timestep conditioning
        timestep = timestep.to(dev).long().expand(B).contiguous() if timestep.dim() == 0 else timestep.to(dev).long()
        tsin = ops.timestep_sinusoid(timestep, 256)
        te = "adaln_single.emb.timestep_embedder."
        t_pre1 = torch.empty(B, D, **b16) if save else None
        t_h = torch.empty(B, D, **b16)
        ops.gemm(tsin, self._w16(te + "linear_1.weight"), bias=self._w32(te + "linear_1.bias"), act=ops.ACT_SILU,
                 pre_out=t_pre1, out_bf16=t_h)
        embedded = torch.empty(B, D, **f32)
        emb_pre = torch.empty(B, D, **b16) if save else None
        semb = torch.empty(B, D, **b16)
        ops.gemm(t_h, self._w16(te + "linear_2.weight"), bias=self._w32(te + "linear_2.bias"), act=ops.ACT_SILU,
                 pre_out=emb_pre, pre_out_f32=embedded, out_bf16=semb)
        tmod = torch.empty(B, 6 * D, **f32)
        ops.gemm(semb, self._w16("adaln_single.linear.weight"), bias=self._w32("adaln_single.linear.bias"), out_f32=tmod)
        # K14 caption projection
        ctx16 = ctx_in.to(dev).reshape(B * L, self.Cc)
        ctx16 = ctx16 if ctx16.dtype == torch.bfloat16 else ctx16.to(torch.bfloat16)
        c_pre = torch.empty(B * L, D, **b16) if save else None
        c1 = torch.empty(B * L, D, **b16)
        ops.gemm(ctx16, self._w16("caption_projection.linear_1.weight"), bias=self._w32("caption_projection.linear_1.bias"),
                 act=ops.ACT_GELU_TANH, pre_out=c_pre, out_bf16=c1)
        ctx = torch.empty(B * L, D, **b16)
        ops.gemm(c1, self._w16("caption_projection.linear_2.weight"), bias=self._w32("caption_projection.linear_2.bias"),
                 out_bf16=ctx)
        if mask is not None:
            ctx_bias = ((1 - mask.to(dev).to(torch.float32)) * -10000.0).contiguous()
        else:
            ctx_bias = None
        # 28 blocks, one C call each
        s = self._shape(B, N, L, lora=not merged)
        ws_bytes = lib.tdm_dit_block_ws_bytes(C.byref(s), int(save))
        xs = [x]
        wss = []
        if not save:
            if self._infer_ws is None or self._infer_ws.numel() < ws_bytes:
                self._infer_ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
        cur = x
        nxt = torch.empty(M, D, **f32)
        for i in range(self.depth):
            if save:
                wsb = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
                wss.append(wsb)
                nxt = torch.empty(M, D, **f32)
            else:
                wsb = self._infer_ws
            check(lib.tdm_dit_block_fwd(C.byref(s), C.byref(bw[i]), ptr(cur), ptr(tmod), ptr(ctx), ptr(ctx_bias),
                                        ptr(nxt), ptr(wsb), int(save), st))
            if save:
                xs.append(nxt)
                cur = nxt
            else:
                cur, nxt = nxt, cur
        # K15 final layer
        modf = ops.modulation(self._w32("scale_shift_table"), embedded, False)
        hf, meanf, rstdf = ops.ln_modulate_fwd(cur, modf[:, 0], modf[:, 1], N, self.eps, want_stats=save)
        y = torch.empty(M, p * p * self.Cout, **f32)
        ops.gemm(hf, self._w16("proj_out.weight"), bias=self._w32("proj_out.bias"), out_f32=y)
        out = ops.unpatchify(y, B, self.Cout, Hh, Ww, p)
        saved = None
        if save:
            saved = _Saved()
            saved.B, saved.N, saved.L, saved.HW, saved.shape = B, N, L, (Hh, Ww), s
            saved.patches, saved.tsin, saved.t_pre1, saved.t_h = patches, tsin, t_pre1, t_h
            saved.emb_pre, saved.semb, saved.tmod, saved.ctx_in = emb_pre, semb, tmod, ctx16
            saved.c_pre, saved.c1, saved.ctx, saved.ctx_bias = c_pre, c1, ctx, ctx_bias
            saved.xs, saved.ws, saved.modf, saved.hf, saved.meanf, saved.rstdf = xs, wss, modf, hf, meanf, rstdf
        return out, saved

    # ------------------------------------------------------------------ backward
    def _lin_wgrad(self, dy16, x16, wname, bname):
        """dW += dy^T x ; db += colsum(dy) into the flat gradient arena."""
        gw = self._g32(wname)
        gw2 = gw.view(gw.shape[0], -1)
        ops.gemm(dy16, x16, a_mn_major=True, b_mn_major=True, res=gw2, out_f32=gw2)
        ops.colsum(dy16, self._g32(bname))

    def backward_impl(self, saved: _Saved, d_out, on_block_done=None):
        """Accumulate parameter gradients for one saved forward, given d(out) for the first d_out.shape[1]
        output channels (the reference only consumes the epsilon half: models.py:40, predictor.py:35).
        `on_block_done(i)` fires after block i's gradients are complete (bucketed all-reduce hook)."""
        lib = _lib.lib()
        st = stream_handle()
        dev = self.device
        D, p = self.D, self.p
        B, N, L = saved.B, saved.N, saved.L
        M = B * N
        tb = self.train_base
        self._ensure_grads()
        if self._block_structs is None:
            self._build_block_structs()
        bw, bg = self._block_structs
        f32 = dict(device=dev, dtype=torch.float32)
        b16 = dict(device=dev, dtype=torch.bfloat16)
        s = saved.shape
        # final layer
        dy = ops.unpatchify_bwd(d_out.to(torch.float32), self.Cout, p)
        if tb:
            self._lin_wgrad(dy, saved.hf, "proj_out.weight", "proj_out.bias")
        dhf = torch.empty(M, D, **b16)
        ops.gemm(dy, self._w16("proj_out.weight"), b_mn_major=True, out_bf16=dhf)
        dmodf = torch.zeros(B, 2, D, **f32) if tb else None
        dx, _ = ops.ln_modulate_bwd(dhf, saved.xs[-1], saved.meanf, saved.rstdf, saved.modf[:, 1], N,
                                    dshift=dmodf[:, 0] if tb else None, dscale=dmodf[:, 1] if tb else None)
        d_emb = torch.zeros(B, D, **f32) if tb else None
        if tb:
            ops.modulation_bwd(dmodf, self._g32("scale_shift_table"), d_emb, False, True)
        # blocks
        sc_bytes = lib.tdm_dit_block_bwd_scratch_bytes(C.byref(s))
        if self._scratch is None or self._scratch.numel() < sc_bytes:
            self._scratch = torch.empty(sc_bytes, device=dev, dtype=torch.uint8)
        d_tmod = torch.zeros(B, 6 * D, **f32) if tb else None
        d_ctx = torch.zeros(B * L, D, **f32) if tb else None
        dx_next = torch.empty(M, D, **f32)
        for i in reversed(range(self.depth)):
            g = bg[i]
            g.d_tmod, g.d_ctx = ptr(d_tmod), ptr(d_ctx)
            check(lib.tdm_dit_block_bwd(C.byref(s), C.byref(bw[i]), C.byref(g), ptr(saved.xs[i]), ptr(saved.ctx),
                                        ptr(saved.ctx_bias), ptr(dx), ptr(dx_next), ptr(saved.ws[i]),
                                        ptr(self._scratch), st))
            dx, dx_next = dx_next, dx
            saved.ws[i] = None  # release this block's activations
            if on_block_done is not None:
                on_block_done(i)
        if tb:
            te = "adaln_single.emb.timestep_embedder."
            # patch embed
            dx16 = ops.cast_bf16(dx)
            self._lin_wgrad(dx16, saved.patches, "pos_embed.proj.weight", "pos_embed.proj.bias")
            # caption projection
            dctx16 = ops.cast_bf16(d_ctx)
            self._lin_wgrad(dctx16, saved.c1, "caption_projection.linear_2.weight", "caption_projection.linear_2.bias")
            dc1 = torch.empty(B * L, D, **b16)
            ops.gemm(dctx16, self._w16("caption_projection.linear_2.weight"), b_mn_major=True,
                     act=ops.ACT_GELU_TANH_BWD, act_in=saved.c_pre, out_bf16=dc1)
            self._lin_wgrad(dc1, saved.ctx_in, "caption_projection.linear_1.weight", "caption_projection.linear_1.bias")
            # timestep conditioning: tmod = linear(silu(embedded)); embedded = linear_2(silu(linear_1(sinusoid)))
            dtmod16 = ops.cast_bf16(d_tmod)
            self._lin_wgrad(dtmod16, saved.semb, "adaln_single.linear.weight", "adaln_single.linear.bias")
            ops.gemm(dtmod16, self._w16("adaln_single.linear.weight"), b_mn_major=True, act=ops.ACT_SILU_BWD,
                     act_in=saved.emb_pre, res=d_emb, out_f32=d_emb)
            demb16 = ops.cast_bf16(d_emb)
            self._lin_wgrad(demb16, saved.t_h, te + "linear_2.weight", te + "linear_2.bias")
            dh16 = torch.empty(B, D, **b16)
            ops.gemm(demb16, self._w16(te + "linear_2.weight"), b_mn_major=True, act=ops.ACT_SILU_BWD,
                     act_in=saved.t_pre1, out_bf16=dh16)
            self._lin_wgrad(dh16, saved.tsin, te + "linear_1.weight", te + "linear_1.bias")
        saved.xs = saved.ws = None
        return None

    def forward(self, hidden_states, timestep=None, added_cond_kwargs=None, encoder_hidden_states=None,
                encoder_attention_mask=None, return_dict=False, **unused):
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if needs_grad:
            out = _DiTFunction.apply(self._anchor, self, hidden_states, timestep, encoder_hidden_states,
                                     encoder_attention_mask)
        else:
            out, _ = self._forward_impl(hidden_states, timestep, encoder_hidden_states, encoder_attention_mask, False)
        if return_dict:
            return {"sample": out}
        return (out,)


class _DiTFunction(torch.autograd.Function):
    """Autograd shim: forward keeps the activations, backward runs the hand-written backward kernels and
    accumulates straight into the parameters' flat .grad arena (inputs get no gradient: the reference
    never differentiates w.r.t. latents or text embeddings)."""

    @staticmethod
    def forward(ctx, anchor, model, latent, timestep, ehs, mask):
        out, saved = model._forward_impl(latent, timestep, ehs, mask, True)
        ctx.model, ctx.saved = model, saved
        return out

    @staticmethod
    def backward(ctx, d_out):
        hook = getattr(ctx.model, "_on_block_done", None)
        ctx.model.backward_impl(ctx.saved, d_out.contiguous(), hook)
        ctx.saved = None
        return None, None, None, None, None, None
