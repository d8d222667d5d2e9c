"""ORACLE (test infrastructure, never on the product path): fp32 PyTorch restatement of the DECODER of
`diffusers.AutoencoderTiny` (TAESD, checkpoint "madebyollin/taesd"), which the reference loads at main.py:166, moves to
the device at main.py:338 and calls at training.py:113-115:

    images = vae.decode(latents.to(vae.dtype) / vae.config.scaling_factor, return_dict=False)[0].clamp(-1, 1) * 0.5 + 0.5

The arithmetic lives in the un-vendored third-party package `diffusers` (only pin: check_min_version("0.26.0.dev0"),
main.py:68; absent from /root/reference, not installable here), so this file restates its published algorithm
(`diffusers/models/autoencoders/vae.py`: `DecoderTiny`, `AutoencoderTinyBlock`; `autoencoder_tiny.py`: config defaults
decoder_block_out_channels (64, 64, 64, 64), num_decoder_blocks (3, 3, 3, 1), act_fn "relu", upsample "nearest",
upsampling_scaling_factor 2, latent_channels 4, out_channels 3, scaling_factor 1.0) with diffusers' parameter names
(`decoder.layers.<i>...`), so a real TAESD `state_dict` loads.  PARITY UNPINNED against diffusers itself: the reference
holds no fixture for the VAE; tests/test_taesd.py arms a comparison with the real class wherever `diffusers` imports.

    DecoderTiny.forward(x):  x = tanh(x / 3) * 3
                             conv3x3(4 -> 64) , ReLU
                             3 x Block(64), Upsample(2, nearest), conv3x3(64 -> 64, no bias)
                             3 x Block(64), Upsample(2, nearest), conv3x3(64 -> 64, no bias)
                             3 x Block(64), Upsample(2, nearest), conv3x3(64 -> 64, no bias)
                             1 x Block(64), conv3x3(64 -> 3)
                             return x * 2 - 1
    Block(x) = ReLU( conv3x3(ReLU(conv3x3(ReLU(conv3x3(x))))) + x )     (skip = Identity: in == out channels)
"""
from __future__ import annotations

import types

import torch
import torch.nn as nn


class AutoencoderTinyBlock(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1), nn.ReLU(), nn.Conv2d(cout, cout, 3, padding=1),
                                  nn.ReLU(), nn.Conv2d(cout, cout, 3, padding=1))
        self.skip = nn.Conv2d(cin, cout, 1, bias=False) if cin != cout else nn.Identity()
        self.fuse = nn.ReLU()

    def forward(self, x):
        return self.fuse(self.conv(x) + self.skip(x))


class DecoderTiny(nn.Module):
    def __init__(self, in_channels=4, out_channels=3, num_blocks=(3, 3, 3, 1), block_out_channels=(64, 64, 64, 64),
                 upsampling_scaling_factor=2):
        super().__init__()
        layers = [nn.Conv2d(in_channels, block_out_channels[0], 3, padding=1), nn.ReLU()]
        for i, nb in enumerate(num_blocks):
            final = i == len(num_blocks) - 1
            c = block_out_channels[i]
            layers += [AutoencoderTinyBlock(c, c) for _ in range(nb)]
            if not final:
                layers.append(nn.Upsample(scale_factor=upsampling_scaling_factor, mode="nearest"))
            layers.append(nn.Conv2d(c, c if not final else out_channels, 3, padding=1, bias=final))
        self.layers = nn.Sequential(*layers)

    def forward(self, x):
        x = torch.tanh(x / 3) * 3
        x = self.layers(x)
        return x.mul(2).sub(1)


class OracleAutoencoderTiny(nn.Module):
    """Decoder half only (the reference never encodes: latents come from the student)."""

    def __init__(self):
        super().__init__()
        self.decoder = DecoderTiny()
        self.config = types.SimpleNamespace(scaling_factor=1.0, latent_channels=4, latent_magnitude=3, latent_shift=0.5)

    @property
    def dtype(self):
        return self.decoder.layers[0].weight.dtype

    def decode(self, x, return_dict=False):
        out = self.decoder(x)
        return {"sample": out} if return_dict else (out,)


def validation_images(vae, latents):
    """training.py:113-117: decode, clamp to [-1, 1], map to [0, 1]."""
    return vae.decode(latents.to(vae.dtype) / vae.config.scaling_factor, return_dict=False)[0].clamp(-1, 1) * 0.5 + 0.5
