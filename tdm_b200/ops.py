"""Thin torch-tensor wrappers over the C ABI (one function per exported kernel family).

These only validate devices/dtypes, pull raw pointers and the current stream, and call the library.
They never compute anything in PyTorch.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, ptr, stream_handle

ACT_NONE, ACT_GELU_TANH, ACT_GELU_TANH_BWD = 0, 1, 2


def _req(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda:
        raise _lib.TdmError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise _lib.TdmError(f"{name} must be {dtype}, got {t.dtype}")
    if t.dim() >= 2 and t.stride(-1) != 1:
        raise _lib.TdmError(f"{name} must have a contiguous last dimension")


def gemm(a, b, *, a_mn_major=False, b_mn_major=False, a2=None, b2=None, seg_n=0, seg_k2=0,
         alpha=1.0, bias=None, act=ACT_NONE, act_in=None, pre_out=None, gate=None, rows_per_gate=1,
         res=None, out_f32=None, out_bf16=None, tile_n=0):
    """C[M,N] = alpha*(A.B^T + A2.B2^T) with the fused epilogue of include/tdm_b200.h (TdmGemm).

    a: [M,K] (or [K,M] when a_mn_major); b: [N,K] (or [K,N] when b_mn_major); 2-D views with unit
    inner stride (row stride = leading dimension)."""
    for t, n in ((a, "a"), (b, "b"), (a2, "a2"), (b2, "b2"), (act_in, "act_in"), (pre_out, "pre_out"),
                 (out_bf16, "out_bf16")):
        _req(t, torch.bfloat16, n)
    for t, n in ((bias, "bias"), (gate, "gate"), (res, "res"), (out_f32, "out_f32")):
        _req(t, torch.float32, n)
    g = _lib.TdmGemm()
    if a_mn_major:
        g.K, g.M = a.shape
    else:
        g.M, g.K = a.shape
    g.N = b.shape[1] if b_mn_major else b.shape[0]
    kb = b.shape[0] if b_mn_major else b.shape[1]
    if kb != g.K:
        raise _lib.TdmError(f"gemm: contraction mismatch {g.K} vs {kb}")
    g.a_mn_major, g.b_mn_major = int(a_mn_major), int(b_mn_major)
    g.A, g.lda = ptr(a), a.stride(0)
    g.B, g.ldb = ptr(b), b.stride(0)
    if a2 is not None:
        g.K2 = b2.shape[0] if b_mn_major else b2.shape[1]
        g.A2, g.lda2 = ptr(a2), a2.stride(0)
        g.B2, g.ldb2 = ptr(b2), b2.stride(0)
        g.seg_n, g.seg_k2 = seg_n, seg_k2
    g.alpha, g.act = float(alpha), int(act)
    g.bias = ptr(bias)
    if act_in is not None:
        g.act_in, g.ld_act = ptr(act_in), act_in.stride(0)
    if pre_out is not None:
        g.pre_out, g.ld_pre = ptr(pre_out), pre_out.stride(0)
    if gate is not None:
        g.gate, g.rows_per_gate, g.gate_stride = ptr(gate), int(rows_per_gate), gate.stride(0)
    if res is not None:
        g.res, g.ld_res = ptr(res), res.stride(0)
    if out_f32 is not None:
        g.out_f32, g.ld_out_f32 = ptr(out_f32), out_f32.stride(0)
    if out_bf16 is not None:
        g.out_bf16, g.ld_out_bf16 = ptr(out_bf16), out_bf16.stride(0)
    g.tile_n = tile_n
    check(_lib.lib().tdm_gemm_bf16(C.byref(g), stream_handle()))
