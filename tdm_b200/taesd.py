"""Drop-in for `diffusers.AutoencoderTiny` as the reference uses it (main.py:166 `from_pretrained("madebyollin/taesd")`,
180 `.requires_grad_(False)`, 338 `.to(device, dtype=)`, training.py:113-115 `.decode(latents / vae.config.scaling_factor,
return_dict=False)[0]`, `.dtype`, `.config.scaling_factor`): the TAESD DECODER on libtdm_b200's kernels
(csrc/taesd.cu: tcgen05 implicit-GEMM 3x3 convolutions over NHWC bf16 activations).  The reference never encodes.

Parameters keep diffusers' names (`decoder.layers.<i>...`), fp32 masters in PyTorch layout; the 64 -> 64 convolutions
are re-packed once to the kernel's bf16 [tap][out][in] layout.  No network here: `from_pretrained` needs a registered
provider (north_star: random-init weights), never a silent download."""
from __future__ import annotations

import ctypes as C
import math
import types

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, ptr, stream_handle

# (layer index in DecoderTiny.layers, kind): diffusers' nn.Sequential numbering (ReLU / Upsample layers hold no parameters)
_BLOCK_LAYERS = [2, 3, 4, 7, 8, 9, 12, 13, 14, 17]
_UP_CONV_LAYERS = [6, 11, 16]
_CONV_IN, _CONV_OUT = 0, 18


class AutoencoderTiny(nn.Module):
    _pretrained_provider = None

    def __init__(self, device="cuda", seed=None, init_weights=True, max_images_per_launch=16):
        super().__init__()
        self.config = types.SimpleNamespace(scaling_factor=1.0, latent_channels=4, latent_magnitude=3, latent_shift=0.5,
                                            decoder_block_out_channels=(64, 64, 64, 64), num_decoder_blocks=(3, 3, 3, 1))
        self.max_images_per_launch = max_images_per_launch
        shapes = {f"decoder.layers.{_CONV_IN}.weight": (64, 4, 3, 3), f"decoder.layers.{_CONV_IN}.bias": (64,)}
        for li in _BLOCK_LAYERS:
            for ci in (0, 2, 4):
                shapes[f"decoder.layers.{li}.conv.{ci}.weight"] = (64, 64, 3, 3)
                shapes[f"decoder.layers.{li}.conv.{ci}.bias"] = (64,)
        for li in _UP_CONV_LAYERS:
            shapes[f"decoder.layers.{li}.weight"] = (64, 64, 3, 3)
        shapes[f"decoder.layers.{_CONV_OUT}.weight"] = (3, 64, 3, 3)
        shapes[f"decoder.layers.{_CONV_OUT}.bias"] = (3,)
        gen = torch.Generator().manual_seed(seed if seed is not None else torch.initial_seed() % (2 ** 63))
        self._names = list(shapes)
        for name, shape in shapes.items():
            t = torch.zeros(shape)
            if init_weights:  # nn.Conv2d default: kaiming_uniform(a = sqrt 5) == U(+-1/sqrt(fan_in)) for weight and bias
                fan_in = (shapes[name.rsplit(".", 1)[0] + ".weight"][1]) * 9
                t = (torch.rand(shape, generator=gen) * 2 - 1) / math.sqrt(fan_in)
            self._set(name, nn.Parameter(t.to(device), requires_grad=False))
        self._packed = None
        self._ws = None

    def _set(self, name, param):
        mod = self
        parts = name.split(".")
        for part in parts[:-1]:
            if part not in mod._modules:
                mod.add_module(part, nn.Module())
            mod = mod._modules[part]
        mod.register_parameter(parts[-1], param)

    # ------------------------------------------------------------------ reference call surface
    @classmethod
    def from_pretrained(cls, path, **kw):
        if cls._pretrained_provider is None:
            raise _lib.TdmError(f"AutoencoderTiny.from_pretrained({path!r}): no weight provider registered (offline build)")
        return cls._pretrained_provider(path, **kw)

    @property
    def dtype(self):
        return torch.float32  # fp32 masters; the kernels compute in bf16 with fp32 accumulation

    @property
    def device(self):
        return next(self.parameters()).device

    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)
        for p in self.parameters():
            if p.dtype != torch.float32:
                raise _lib.TdmError("AutoencoderTiny keeps fp32 master weights (compute runs in bf16 internally)")
        self._packed = self._ws = None
        return self

    def load_state_dict(self, state_dict, strict=True, **kw):
        out = super().load_state_dict({k: v for k, v in state_dict.items() if k.startswith("decoder.")}, strict=strict, **kw)
        self._packed = None
        return out

    # ------------------------------------------------------------------ weights for the kernels
    def _pack(self):
        sd = dict(self.named_parameters())
        dev = self.device
        if not dev.type == "cuda":
            raise _lib.TdmError("AutoencoderTiny lives on a CUDA device (no CPU fallback)")

        def taps(w):  # [out, in, ky, kx] fp32 -> [tap = ky*3+kx][out][in] bf16
            return w.detach().permute(2, 3, 0, 1).reshape(9 * 64, 64).to(torch.bfloat16).contiguous()

        keep = []
        w = _lib.TdmTaesdWeights()
        f32 = lambda n: sd[n].detach().float().contiguous()
        t = f32(f"decoder.layers.{_CONV_IN}.weight"); keep.append(t); w.conv_in_w = ptr(t)
        t = f32(f"decoder.layers.{_CONV_IN}.bias"); keep.append(t); w.conv_in_b = ptr(t)
        i = 0
        for li in _BLOCK_LAYERS:
            for ci in (0, 2, 4):
                tw = taps(sd[f"decoder.layers.{li}.conv.{ci}.weight"]); keep.append(tw)
                tb = f32(f"decoder.layers.{li}.conv.{ci}.bias"); keep.append(tb)
                w.block_w[i], w.block_b[i] = ptr(tw), ptr(tb)
                i += 1
        for j, li in enumerate(_UP_CONV_LAYERS):
            tw = taps(sd[f"decoder.layers.{li}.weight"]); keep.append(tw)
            w.up_w[j] = ptr(tw)
        t = f32(f"decoder.layers.{_CONV_OUT}.weight"); keep.append(t); w.conv_out_w = ptr(t)
        t = f32(f"decoder.layers.{_CONV_OUT}.bias"); keep.append(t); w.conv_out_b = ptr(t)
        self._packed = (w, keep)

    # ------------------------------------------------------------------ decode
    def decode(self, x, return_dict=False, to_image=False):
        """latents [B,4,H,W] (fp32 / fp16 / bf16, CUDA) -> [B,3,8H,8W] fp32; `to_image` fuses training.py:113-117's
        clamp(-1, 1) * 0.5 + 0.5 into the last kernel."""
        if not x.is_cuda:
            raise _lib.TdmError("AutoencoderTiny.decode: CUDA tensors only (no CPU fallback)")
        if self._packed is None:
            self._pack()
        w, _ = self._packed
        x = x.contiguous()
        B, Cl, H, W = x.shape
        if Cl != 4:
            raise _lib.TdmError("AutoencoderTiny.decode expects 4 latent channels")
        out = torch.empty(B, 3, 8 * H, 8 * W, device=x.device, dtype=torch.float32)
        lib = _lib.lib()
        step = max(1, min(B, self.max_images_per_launch))
        need = lib.tdm_taesd_ws_bytes(step, H, W)
        if self._ws is None or self._ws.numel() < need or self._ws.device != x.device:
            self._ws = torch.empty(need, device=x.device, dtype=torch.uint8)
        for b0 in range(0, B, step):
            n = min(step, B - b0)
            check(lib.tdm_taesd_decode(C.byref(w), ptr(x[b0:b0 + n]), _lib.dtype_enum(x), n, H, W, ptr(out[b0:b0 + n]), int(to_image),
                                       ptr(self._ws), self._ws.numel(), stream_handle()))
        if return_dict:
            return {"sample": out}
        return (out,)


def decode_to_images(vae, latents):
    """training.py:113-117: vae.decode(latents.to(vae.dtype) / vae.config.scaling_factor)[0].clamp(-1, 1) * 0.5 + 0.5, the
    scaling (1.0 for TAESD) applied on the way in, the clamp / affine fused into the decoder's last kernel."""
    sf = vae.config.scaling_factor
    return vae.decode(latents if sf == 1.0 else latents / sf, return_dict=False, to_image=True)[0]
