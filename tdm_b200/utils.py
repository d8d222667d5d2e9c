"""Mirror of the reference's src/utils.py hot-path functions (same names, argument meaning and error
behaviour) running on the library's fused kernels."""
from __future__ import annotations

import torch

from . import ops


def extract_into_tensor(a, t, x_shape):
    """utils.py:70-73 (index plumbing on a [B] tensor)."""
    b, *_ = t.shape
    out = a.gather(-1, t)
    return out.reshape(b, *((1,) * (len(x_shape) - 1)))


def predicted_origin(model_output, timesteps, sample, prediction_type, alphas, sigmas):
    """utils.py:47-59 as one kernel.  model_output: the epsilon (or v) prediction, fp32 [B,C,H,W]."""
    if prediction_type not in ops.PRED_TYPES:
        raise ValueError(f"Prediction type {prediction_type} currently not supported.")
    _, x0, _ = ops.solver_step(model_output.float(), sample, timesteps.long(), alphas, sigmas, prediction_type)
    return x0


def append_dims(x, target_dims):
    """utils.py:62-67."""
    dims_to_append = target_dims - x.ndim
    if dims_to_append < 0:
        raise ValueError(f"input has {x.ndim} dims but target_dims is {target_dims}, which is less")
    return x[(...,) + (None,) * dims_to_append]


def get_module_kohya_state_dict(module, prefix: str, dtype: torch.dtype, adapter_name: str = "default"):
    """utils.py:76-93: LoRA weights in the kohya-ss naming that `pipe.load_lora_weights` consumes (README.md:49, 96, 119):
    `base_model.model` -> prefix, lora_A -> lora_down, lora_B -> lora_up, every dot but the last two -> underscore, plus
    one `.alpha` entry per adapted linear.  `module` is a tdm_b200 Transformer2DModel with add_lora() applied (it
    exposes `peft_state_dict()` / `peft_config` the way a peft-wrapped model would through get_peft_model_state_dict)."""
    kohya_ss_state_dict = {}
    for peft_key, weight in module.peft_state_dict(adapter_name=adapter_name).items():
        if "base_model.model" in peft_key:
            kohya_key = peft_key.replace("base_model.model", prefix)
        else:
            kohya_key = prefix + "." + peft_key
        kohya_key = kohya_key.replace("lora_A", "lora_down")
        kohya_key = kohya_key.replace("lora_B", "lora_up")
        kohya_key = kohya_key.replace(".", "_", kohya_key.count(".") - 2)
        kohya_ss_state_dict[kohya_key] = weight.to(dtype)
        if "lora_down" in kohya_key:
            alpha_key = f'{kohya_key.split(".")[0]}.alpha'
            kohya_ss_state_dict[alpha_key] = torch.tensor(module.peft_config[adapter_name].lora_alpha).to(dtype)
    return kohya_ss_state_dict
