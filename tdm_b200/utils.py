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
