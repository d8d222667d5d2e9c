"""ctypes binding of libtdm_b200.so (the C ABI declared in include/tdm_b200.h).

The product path has NO fallback: if the shared library is missing or a call fails, a RuntimeError
is raised.  Nothing in this package imports `oracle/`.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libtdm_b200.so"

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
        ("gate", vp), ("rows_per_gate", i32), ("gate_stride", i64),
        ("res", vp), ("ld_res", i64),
        ("out_f32", vp), ("ld_out_f32", i64),
        ("out_bf16", vp), ("ld_out_bf16", i64),
        ("tile_n", i32),
    ]


# name -> (restype, argtypes); every symbol include/tdm_b200.h declares must appear here
# (tests/test_abi.py cross-checks this table against the header and the built library).
SIGNATURES = {
    "tdm_last_error": (C.c_char_p, []),
    "tdm_version": (C.c_int, []),
    "tdm_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "tdm_gemm_bf16": (C.c_int, [C.POINTER(TdmGemm), vp]),
}


def _declare(l: C.CDLL) -> None:
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(l, name)
        fn.restype = res
        fn.argtypes = args
