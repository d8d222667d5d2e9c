"""Mirror of the reference's src/utils.py hot-path functions (same names, argument meaning and error
behaviour) running on the library's fused kernels.

`compute_snr` and `append_dims` (utils.py:21-44, 62-67) are dead code in the reference and have
no counterpart here.  `get_module_kohya_state_dict` keeps the reference's name and call signature but is written from
this package's own LoRA name table (checkpoint.kohya_lora_state_dict)."""
from __future__ import annotations

import torch

from . import ops


def extract_into_tensor(a, t, x_shape):
    """utils.py:70-73: a[t] broadcastable against a tensor of shape x_shape -> [B, 1, 1, ...] in a's dtype."""
    trailing = (1,) * (len(x_shape) - 1)
    return torch.gather(a, -1, t).view(t.shape[0], *trailing)


def predicted_origin(model_output, timesteps, sample, prediction_type, alphas, sigmas):
    """utils.py:47-59 as one kernel.

    Bit-identical to the reference for fp32 model outputs -- the only case on the hot path (accelerate hands autocast
    outputs back as fp32; models.py:46, predictor.py:43-56, main.py:484-490) -- and returned in fp32 as torch's type
    promotion gives there.  A lower-precision `model_output` is evaluated in fp32 as well and the result cast to the
    dtype promotion would produce; the reference's eager sequence rounds after every op in that case, so only the
    dtype, not the last bit, is mirrored."""
    if prediction_type not in ops.PRED_TYPES:
        raise ValueError(f"Prediction type {prediction_type} currently not supported.")
    promoted = torch.promote_types(torch.promote_types(model_output.dtype, sample.dtype), alphas.dtype)
    mo = model_output if model_output.dtype == torch.float32 else model_output.float()
    _, x0, _ = ops.solver_step(mo, sample, timesteps.long(), alphas, sigmas, prediction_type)
    return x0 if promoted == torch.float32 else x0.to(promoted)


def get_module_kohya_state_dict(module, prefix: str, dtype: torch.dtype, adapter_name: str = "default"):
    """utils.py:76-93 call surface: the LoRA adapters of `module` (a Transformer2DModel after add_lora(), or anything
    exposing `peft_state_dict()` / `peft_config`) in the kohya-ss naming `pipe.load_lora_weights` reads
    (README.md:49, 96, 119)."""
    from .checkpoint import kohya_lora_state_dict
    return kohya_lora_state_dict(module.peft_state_dict(adapter_name=adapter_name), prefix, dtype,
                                 module.peft_config[adapter_name].lora_alpha)
