"""ctypes binding of libtdm_b200.so (the C ABI declared in include/tdm_b200.h).

The product path has NO fallback: if the shared library is missing or a call fails, a RuntimeError
is raised.  Nothing in this package imports `oracle/`.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_HERE = Path(__file__).resolve().parent
import os as _os

LIB_PATH = Path(_os.environ["TDM_B200_LIB"]) if _os.environ.get("TDM_B200_LIB") else _HERE / "libtdm_b200.so"

_lib = None


class TdmError(RuntimeError):
    pass


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise TdmError(
                f"{LIB_PATH} is missing: build it with `python -m tdm_b200.build` "
                "(there is no CPU / PyTorch fallback for the TDM hot path)"
            )
        _lib = C.CDLL(str(LIB_PATH))
        _lib.tdm_last_error.restype = C.c_char_p
        _declare(_lib)
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise TdmError(f"libtdm_b200 error {rc}: {lib().tdm_last_error().decode(errors='replace')}")


def ptr(t) -> int:
    """Device pointer of a torch tensor (or None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def stream_handle() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream


vp = C.c_void_p
i32 = C.c_int32
i64 = C.c_int64
f32 = C.c_float


class TdmGemm(C.Structure):
    _fields_ = [
        ("M", i32), ("N", i32), ("K", i32),
        ("a_mn_major", i32), ("b_mn_major", i32),
        ("A", vp), ("lda", i64),
        ("B", vp), ("ldb", i64),
        ("K2", i32), ("seg_n", i32), ("seg_k2", i32),
        ("A2", vp), ("lda2", i64),
        ("B2", vp), ("ldb2", i64),
        ("alpha", f32), ("act", i32),
        ("bias", vp),
        ("act_in", vp), ("ld_act", i64),
        ("pre_out", vp), ("ld_pre", i64),
        ("pre_out_f32", vp), ("ld_pre_f32", i64),
        ("gate", vp), ("rows_per_gate", i32), ("gate_stride", i64),
        ("res", vp), ("ld_res", i64), ("res_row_mod", i32),
        ("out_f32", vp), ("ld_out_f32", i64),
        ("out_bf16", vp), ("ld_out_bf16", i64),
        ("tile_n", i32), ("split_k", i32), ("cta_pair", i32), ("seg_k", i32), ("seg_m", i32), ("seg_nb", i32), ("epilogue", i32),
    ]


class TdmAttn(C.Structure):
    _fields_ = [
        ("B", i32), ("H", i32), ("Sq", i32), ("Sk", i32), ("d", i32),
        ("scale", f32),
        ("Q", vp), ("q_stride_b", i64), ("q_stride_s", i64), ("q_stride_h", i64),
        ("K", vp), ("k_stride_b", i64), ("k_stride_s", i64), ("k_stride_h", i64),
        ("V", vp), ("v_stride_b", i64), ("v_stride_s", i64), ("v_stride_h", i64),
        ("bias", vp),
        ("O", vp), ("o_stride_b", i64), ("o_stride_s", i64), ("o_stride_h", i64),
        ("lse", vp),
        ("dO", vp), ("delta", vp), ("dq_accum", vp),
        ("dQ", vp), ("dq_stride_b", i64), ("dq_stride_s", i64), ("dq_stride_h", i64),
        ("dK", vp), ("dk_stride_b", i64), ("dk_stride_s", i64), ("dk_stride_h", i64),
        ("dV", vp), ("dv_stride_b", i64), ("dv_stride_s", i64), ("dv_stride_h", i64),
    ]


class TdmDitShape(C.Structure):
    _fields_ = [("B", i32), ("N", i32), ("D", i32), ("H", i32), ("L", i32), ("Dff", i32),
                ("lora_r", i32), ("lora_scale", f32), ("eps", f32), ("train_base", i32)]


_BLOCK_BASE = ["table", "w_qkv", "b_qkv", "w_o", "b_o", "w_q2", "b_q2", "w_kv2", "b_kv2", "w_o2", "b_o2",
               "w_ff1", "b_ff1", "w_ff2", "b_ff2"]
_BLOCK_LORA = [f"lora_{ab}_{n}" for n in ("qkv", "o", "q2", "kv2", "o2", "ff1", "ff2") for ab in ("a", "b")]


class TdmBlockWeights(C.Structure):
    _fields_ = [(n, vp) for n in _BLOCK_BASE + _BLOCK_LORA]


class TdmBlockGrads(C.Structure):
    _fields_ = [(n, vp) for n in _BLOCK_BASE + _BLOCK_LORA + ["d_tmod", "d_ctx"]]


class TdmTaesdWeights(C.Structure):
    _fields_ = [("conv_in_w", vp), ("conv_in_b", vp), ("block_w", vp * 30), ("block_b", vp * 30), ("up_w", vp * 3),
                ("conv_out_w", vp), ("conv_out_b", vp)]


ci, cl, cd, cf = C.c_int, C.c_long, C.c_double, C.c_float

# name -> (restype, argtypes); every symbol include/tdm_b200.h declares must appear here
# (tests/test_abi.py cross-checks this table against the header and the built library).
SIGNATURES = {
    "tdm_last_error": (C.c_char_p, []),
    "tdm_version": (ci, []),
    "tdm_device_info": (ci, [C.POINTER(ci)] * 3),
    "tdm_launch_count": (C.c_uint64, [ci]),
    "tdm_gemm_bf16": (ci, [C.POINTER(TdmGemm), vp]),
    "tdm_attn_fwd": (ci, [C.POINTER(TdmAttn), vp]),
    "tdm_attn_bwd": (ci, [C.POINTER(TdmAttn), vp]),
    "tdm_solver_step": (ci, [vp, vp, cd, vp, ci, vp, vp, vp, vp, ci, ci, vp, ci, vp, vp, vp, ci, ci, ci, ci, ci, vp]),
    "tdm_renoise": (ci, [vp, vp, ci, vp, vp, ci, vp, vp, vp, vp, ci, ci, vp, ci, ci, ci, cl, vp]),
    "tdm_loss_fwd_bwd": (ci, [vp, vp, vp, cd, ci, cd, vp, vp, vp, cf, ci, cl, vp]),
    "tdm_ln_modulate_fwd": (ci, [vp, vp, vp, i64, ci, vp, vp, vp, ci, ci, cf, vp]),
    "tdm_ln_modulate_bwd": (ci, [vp, vp, vp, vp, vp, i64, ci, vp, vp, vp, vp, vp, i64, ci, ci, vp]),
    "tdm_gate_bwd": (ci, [vp, vp, vp, i64, ci, vp, vp, i64, ci, ci, vp]),
    "tdm_colsum_bf16": (ci, [vp, i64, ci, ci, vp, vp]),
    "tdm_modulation": (ci, [vp, vp, ci, vp, ci, ci, ci, vp]),
    "tdm_modulation_bwd": (ci, [vp, vp, vp, ci, ci, ci, ci, ci, vp]),
    "tdm_timestep_sinusoid": (ci, [vp, vp, ci, ci, vp]),
    "tdm_patchify": (ci, [vp, ci, vp, ci, ci, ci, ci, ci, vp]),
    "tdm_unpatchify": (ci, [vp, i64, vp, ci, ci, ci, ci, ci, vp]),
    "tdm_unpatchify_bwd": (ci, [vp, vp, ci, ci, ci, ci, ci, ci, vp]),
    "tdm_cast_f32_bf16": (ci, [vp, vp, i64, vp]),
    "tdm_scheduler_add_noise": (ci, [vp, ci, vp, ci, vp, vp, ci, vp, ci, ci, cl, vp]),
    "tdm_solver_step_bwd": (ci, [vp, vp, cd, ci, vp, vp, vp, ci, ci, vp, vp, ci, ci, ci, ci, vp]),
    "tdm_select_states": (ci, [vp, vp, ci, vp, vp, ci, ci, cl, vp]),
    "tdm_mse_fwd_bwd": (ci, [vp, vp, vp, vp, cf, ci, ci, ci, ci, vp]),
    "tdm_sumsq_f32": (ci, [vp, i64, vp, ci, vp]),
    "tdm_clip_coef": (ci, [vp, cf, cf, vp, vp, vp]),
    "tdm_adamw_step": (ci, [vp, vp, vp, vp, vp, i64, cd, cd, cd, cd, cd, ci, vp, vp]),
    "tdm_ema_update": (ci, [vp, vp, i64, cf, vp]),
    "tdm_lora_merge": (ci, [vp, ci, ci, cf, vp]),
    "tdm_taesd_ws_bytes": (i64, [ci, ci, ci]),
    "tdm_taesd_decode": (ci, [C.POINTER(TdmTaesdWeights), vp, ci, ci, ci, ci, vp, ci, vp, i64, vp]),
    "tdm_dit_block_ws_bytes": (i64, [C.POINTER(TdmDitShape), ci]),
    "tdm_dit_block_bwd_scratch_bytes": (i64, [C.POINTER(TdmDitShape)]),
    "tdm_dit_block_fwd": (ci, [C.POINTER(TdmDitShape), C.POINTER(TdmBlockWeights), vp, vp, vp, vp, vp, vp, ci, vp]),
    "tdm_dit_block_bwd": (ci, [C.POINTER(TdmDitShape), C.POINTER(TdmBlockWeights), C.POINTER(TdmBlockGrads),
                               vp, vp, vp, vp, vp, vp, vp, vp]),
}

DTYPE_ENUM = {"torch.float32": 0, "torch.float16": 1, "torch.bfloat16": 2}


def dtype_enum(t) -> int:
    try:
        return DTYPE_ENUM[str(t.dtype)]
    except KeyError:
        raise TdmError(f"unsupported dtype {t.dtype} (fp32 / fp16 / bf16 only)") from None


def _declare(l: C.CDLL) -> None:
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(l, name)
        fn.restype = res
        fn.argtypes = args
