"""Thin torch-tensor wrappers over the C ABI (one function per exported kernel family).

These only validate devices/dtypes, pull raw pointers and the current stream, and call the library.
They never compute anything in PyTorch.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, ptr, stream_handle

ACT_NONE, ACT_GELU_TANH, ACT_GELU_TANH_BWD, ACT_SILU, ACT_SILU_BWD = 0, 1, 2, 3, 4


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
         alpha=1.0, bias=None, act=ACT_NONE, act_in=None, pre_out=None, pre_out_f32=None, gate=None,
         rows_per_gate=1, res=None, res_row_mod=0, out_f32=None, out_bf16=None, tile_n=0, split_k=0, cta_pair=0, epilogue=0, seg_k=0, seg_m=0, seg_nb=0):
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
    if pre_out_f32 is not None:
        _req(pre_out_f32, torch.float32, "pre_out_f32")
        g.pre_out_f32, g.ld_pre_f32 = ptr(pre_out_f32), pre_out_f32.stride(0)
    if res is not None:
        g.res, g.ld_res, g.res_row_mod = ptr(res), res.stride(0), int(res_row_mod)
    if out_f32 is not None:
        g.out_f32, g.ld_out_f32 = ptr(out_f32), out_f32.stride(0)
    if out_bf16 is not None:
        g.out_bf16, g.ld_out_bf16 = ptr(out_bf16), out_bf16.stride(0)
    g.tile_n = tile_n
    g.split_k = split_k
    g.cta_pair = cta_pair
    g.epilogue = epilogue
    if seg_k or seg_m:  # stacked problems: the tensors span every segment, the descriptor holds ONE problem's K / N
        g.seg_n, g.seg_k, g.seg_m, g.seg_nb = int(seg_n), int(seg_k), int(seg_m), int(seg_nb)
        if seg_k:
            nseg = g.K // int(seg_k)
            g.K, g.N = int(seg_k), nseg * int(seg_n)
        if seg_m:
            g.N = int(seg_nb)
    check(_lib.lib().tdm_gemm_bf16(C.byref(g), stream_handle()))


def _cuda(t, name):
    if t is not None and not t.is_cuda:
        raise _lib.TdmError(f"{name} must be a CUDA tensor (no CPU fallback)")


def _bshd(t, name):
    """(ptr, stride_b, stride_s, stride_h) of a bf16 [B, S, H, d] view with unit stride on d."""
    _req(t, torch.bfloat16, name)
    if t.dim() != 4:
        raise _lib.TdmError(f"{name} must be a [B, S, H, d] view")
    return ptr(t), t.stride(0), t.stride(1), t.stride(2)


def _attn_desc(q, k, v, o, scale, bias, lse):
    a = _lib.TdmAttn()
    a.B, a.Sq, a.H, a.d = q.shape
    a.Sk = k.shape[1]
    a.scale = float(scale)
    a.Q, a.q_stride_b, a.q_stride_s, a.q_stride_h = _bshd(q, "q")
    a.K, a.k_stride_b, a.k_stride_s, a.k_stride_h = _bshd(k, "k")
    a.V, a.v_stride_b, a.v_stride_s, a.v_stride_h = _bshd(v, "v")
    a.O, a.o_stride_b, a.o_stride_s, a.o_stride_h = _bshd(o, "o")
    _req(bias, torch.float32, "bias")
    _req(lse, torch.float32, "lse")
    a.bias, a.lse = ptr(bias), ptr(lse)
    return a


def attn_fwd(q, k, v, o, scale, bias=None, lse=None):
    """o = softmax(q k^T * scale + bias[b, key]) v ; q,k,v,o: bf16 [B,S,H,d] views; lse fp32 [B,H,Sq]."""
    a = _attn_desc(q, k, v, o, scale, bias, lse)
    check(_lib.lib().tdm_attn_fwd(C.byref(a), stream_handle()))


def attn_bwd(q, k, v, o, do, lse, scale, dq, dk, dv, bias=None, delta=None, dq_accum=None):
    a = _attn_desc(q, k, v, o, scale, bias, lse)
    B, Sq, H, d = q.shape
    if delta is None:
        delta = torch.empty(B, H, Sq, device=q.device, dtype=torch.float32)
    if dq_accum is None:
        dq_accum = torch.empty(B * Sq, H * d, device=q.device, dtype=torch.float32)
    if do.stride() != o.stride():
        raise _lib.TdmError("attn_bwd: do must have the same strides as o")
    _req(do, torch.bfloat16, "do")
    a.dO, a.delta, a.dq_accum = ptr(do), ptr(delta), ptr(dq_accum)
    a.dQ, a.dq_stride_b, a.dq_stride_s, a.dq_stride_h = _bshd(dq, "dq")
    a.dK, a.dk_stride_b, a.dk_stride_s, a.dk_stride_h = _bshd(dk, "dk")
    a.dV, a.dv_stride_b, a.dv_stride_s, a.dv_stride_h = _bshd(dv, "dv")
    check(_lib.lib().tdm_attn_bwd(C.byref(a), stream_handle()))


PRED_TYPES = {"epsilon": 0, "v_prediction": 1}


def solver_step(model_out, x_t, t, alphas, sigmas, prediction_type="epsilon", model_out_uncond=None, cfg=1.0,
                t_next=None, alphas_cumprod=None, next_dtype=None, want_eps=False, want_x0=True):
    """Fused eps-select + CFG + predicted_origin (+ scheduler.add_noise(...).to(next_dtype)).
    Returns (eps or None, x0 or None, x_next or None)."""
    if prediction_type not in PRED_TYPES:
        raise ValueError(f"Prediction type {prediction_type} currently not supported.")  # utils.py:56-57
    _req(model_out, torch.float32, "model_out")
    _req(model_out_uncond, torch.float32, "model_out_uncond")
    for x, n in ((x_t, "x_t"), (t, "t"), (alphas, "alphas"), (sigmas, "sigmas"), (t_next, "t_next"),
                 (alphas_cumprod, "alphas_cumprod")):
        _cuda(x, n)
    B, Cout = model_out.shape[0], model_out.shape[1]
    C_ = x_t.shape[1]
    HW = x_t[0, 0].numel()
    model_out, x_t = model_out.contiguous(), x_t.contiguous()
    if model_out_uncond is not None:
        model_out_uncond = model_out_uncond.contiguous()
    eps = torch.empty(x_t.shape, device=x_t.device, dtype=torch.float32) if want_eps else None
    x0 = torch.empty(x_t.shape, device=x_t.device, dtype=torch.float32) if want_x0 else None
    x_next = None
    if t_next is not None:
        _req(alphas_cumprod, torch.float32, "alphas_cumprod")
        x_next = torch.empty(x_t.shape, device=x_t.device, dtype=next_dtype or x_t.dtype)
    if t.dtype != torch.int64 or (t_next is not None and t_next.dtype != torch.int64):
        raise _lib.TdmError("timesteps must be int64")
    if alphas.dtype != sigmas.dtype:
        raise _lib.TdmError("alphas/sigmas dtype mismatch")
    check(_lib.lib().tdm_solver_step(
        ptr(model_out), ptr(model_out_uncond), float(cfg), ptr(x_t), _lib.dtype_enum(x_t), ptr(t), ptr(t_next),
        ptr(alphas), ptr(sigmas), _lib.dtype_enum(alphas), alphas.numel(), ptr(alphas_cumprod),
        PRED_TYPES[prediction_type], ptr(eps), ptr(x0), ptr(x_next),
        _lib.dtype_enum(x_next) if x_next is not None else 0, B, C_, Cout, HW, stream_handle()))
    return eps, x0, x_next


def renoise(states, noise, t1, t2, alphas, sigmas, out_dtype, index=None, mixed=False):
    """Predictor.add_noise / obtain_mixed_noise; `states` is one tensor or a list selected per sample by index."""
    if not isinstance(states, (list, tuple)):
        states = [states]
    states = [s.contiguous() for s in states]
    noise = noise.contiguous()
    for s in states:
        _cuda(s, "states")
    B = states[0].shape[0]
    per = states[0][0].numel()
    out = torch.empty(states[0].shape, device=noise.device, dtype=out_dtype)
    arr_p = (C.c_void_p * len(states))(*[ptr(s) for s in states])
    arr_d = (C.c_int * len(states))(*[_lib.dtype_enum(s) for s in states])
    check(_lib.lib().tdm_renoise(
        C.cast(arr_p, C.c_void_p), C.cast(arr_d, C.c_void_p), len(states), ptr(index), ptr(noise),
        _lib.dtype_enum(noise), ptr(t1), ptr(t2), ptr(alphas), ptr(sigmas), _lib.dtype_enum(alphas), alphas.numel(),
        ptr(out), _lib.dtype_enum(out), int(mixed), B, per, stream_handle()))
    return out


def tdm_loss(model_latents, sd_latents, sd_latents_uncond, cfg, use_huber, huber_c=1e-3, want_grad=True,
             grad_scale=1.0):
    """TDM loss (main.py:514-529) -> (loss fp64 0-dim tensor, d loss / d model_latents fp32 or None)."""
    for x, n in ((model_latents, "model_latents"), (sd_latents, "sd_latents"), (sd_latents_uncond, "sd_latents_uncond")):
        _req(x, torch.float32, n)
    m, s, u = model_latents.contiguous(), sd_latents.contiguous(), sd_latents_uncond.contiguous()
    B = m.shape[0]
    per = m[0].numel()
    scratch = torch.empty(B + 1, device=m.device, dtype=torch.float64)
    grad = torch.empty_like(m) if want_grad else None
    check(_lib.lib().tdm_loss_fwd_bwd(ptr(m), ptr(s), ptr(u), float(cfg), int(bool(use_huber)), float(huber_c),
                                      scratch.data_ptr(), scratch.data_ptr() + 8 * B, ptr(grad), float(grad_scale),
                                      B, per, stream_handle()))
    return scratch[B], grad


def ln_modulate_fwd(x, shift, scale, rows_per_sample, eps=1e-6, want_stats=True):
    """x fp32 [M,D]; shift/scale: fp32 [B,D] views (row stride = mod stride). -> (h bf16, mean, rstd)"""
    _req(x, torch.float32, "x"); _req(shift, torch.float32, "shift"); _req(scale, torch.float32, "scale")
    M, D = x.shape
    if shift.stride(0) != scale.stride(0):
        raise _lib.TdmError("shift/scale must share their sample stride")
    out = torch.empty(M, D, device=x.device, dtype=torch.bfloat16)
    mean = torch.empty(M, device=x.device, dtype=torch.float32) if want_stats else None
    rstd = torch.empty(M, device=x.device, dtype=torch.float32) if want_stats else None
    check(_lib.lib().tdm_ln_modulate_fwd(ptr(x), ptr(shift), ptr(scale), shift.stride(0), rows_per_sample, ptr(out),
                                         ptr(mean), ptr(rstd), M, D, float(eps), stream_handle()))
    return out, mean, rstd


def ln_modulate_bwd(dh, x, mean, rstd, scale, rows_per_sample, dx_in=None, want_bf16=False, dshift=None, dscale=None):
    _req(dh, torch.bfloat16, "dh"); _req(x, torch.float32, "x")
    M, D = x.shape
    dx = torch.empty(M, D, device=x.device, dtype=torch.float32)
    dxb = torch.empty(M, D, device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    check(_lib.lib().tdm_ln_modulate_bwd(ptr(dh), ptr(x), ptr(mean), ptr(rstd), ptr(scale), scale.stride(0),
                                         rows_per_sample, ptr(dx_in), ptr(dx), ptr(dxb), ptr(dshift), ptr(dscale),
                                         dshift.stride(0) if dshift is not None else 0, M, D, stream_handle()))
    return dx, dxb


def gate_bwd(dx, gate, rows_per_sample, y=None, dgate=None):
    _req(dx, torch.float32, "dx"); _req(gate, torch.float32, "gate"); _req(y, torch.bfloat16, "y")
    M, D = dx.shape
    dy = torch.empty(M, D, device=dx.device, dtype=torch.bfloat16)
    check(_lib.lib().tdm_gate_bwd(ptr(dx), ptr(y), ptr(gate), gate.stride(0), rows_per_sample, ptr(dy), ptr(dgate),
                                  dgate.stride(0) if dgate is not None else 0, M, D, stream_handle()))
    return dy


def colsum(a, out):
    _req(a, torch.bfloat16, "a"); _req(out, torch.float32, "out")
    check(_lib.lib().tdm_colsum_bf16(ptr(a), a.stride(0), a.shape[0], a.shape[1], ptr(out), stream_handle()))


def modulation(table, tvec, per_j):
    J, D = table.shape
    B = tvec.shape[0]
    out = torch.empty(B, J, D, device=table.device, dtype=torch.float32)
    check(_lib.lib().tdm_modulation(ptr(table), ptr(tvec.contiguous()), int(per_j), ptr(out), B, J, D, stream_handle()))
    return out


def modulation_bwd(dmod, dtable, dtvec, per_j, accumulate_tvec):
    B, J, D = dmod.shape
    check(_lib.lib().tdm_modulation_bwd(ptr(dmod), ptr(dtable), ptr(dtvec), int(per_j), int(accumulate_tvec), B, J, D,
                                        stream_handle()))


def timestep_sinusoid(t, dim=256):
    out = torch.empty(t.shape[0], dim, device=t.device, dtype=torch.bfloat16)
    check(_lib.lib().tdm_timestep_sinusoid(ptr(t), ptr(out), t.shape[0], dim, stream_handle()))
    return out


def patchify(latent, p):
    latent = latent.contiguous()
    B, C_, H, W = latent.shape
    out = torch.empty(B * (H // p) * (W // p), C_ * p * p, device=latent.device, dtype=torch.bfloat16)
    check(_lib.lib().tdm_patchify(ptr(latent), _lib.dtype_enum(latent), ptr(out), B, C_, H, W, p, stream_handle()))
    return out


def unpatchify(y, B, Cout, H, W, p):
    _req(y, torch.float32, "y")
    out = torch.empty(B, Cout, H, W, device=y.device, dtype=torch.float32)
    check(_lib.lib().tdm_unpatchify(ptr(y), y.stride(0), ptr(out), B, Cout, H, W, p, stream_handle()))
    return out


def unpatchify_bwd(dout, Cout, p):
    dout = dout.contiguous()
    _req(dout, torch.float32, "dout")
    B, Cg, H, W = dout.shape
    dy = torch.empty(B * (H // p) * (W // p), p * p * Cout, device=dout.device, dtype=torch.bfloat16)
    check(_lib.lib().tdm_unpatchify_bwd(ptr(dout), ptr(dy), B, Cg, Cout, H, W, p, stream_handle()))
    return dy


def cast_bf16(x):
    _req(x, torch.float32, "x")
    x = x.contiguous()
    out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    check(_lib.lib().tdm_cast_f32_bf16(ptr(x), ptr(out), x.numel(), stream_handle()))
    return out


def select_states(states, index, out_dtype=None):
    """out[b] = states[index[b]][b] (main.py:479 as a per-sample gather)."""
    states = [s.contiguous() for s in states]
    for s in states:
        _cuda(s, "states")
    B = states[0].shape[0]
    per = states[0][0].numel()
    out = torch.empty(states[0].shape, device=states[0].device, dtype=out_dtype or states[0].dtype)
    arr_p = (C.c_void_p * len(states))(*[ptr(s) for s in states])
    arr_d = (C.c_int * len(states))(*[_lib.dtype_enum(s) for s in states])
    if index.dtype != torch.int64:
        raise _lib.TdmError("index must be int64")
    check(_lib.lib().tdm_select_states(C.cast(arr_p, C.c_void_p), C.cast(arr_d, C.c_void_p), len(states), ptr(index),
                                       ptr(out), _lib.dtype_enum(out), B, per, stream_handle()))
    return out


def mse_loss(pred, target, want_grad=True, grad_scale=1.0):
    """mean((pred[:, :C] - target)^2) -> (fp64 0-dim loss, grad [B,C,H,W] fp32 or None)."""
    _req(pred, torch.float32, "pred"); _req(target, torch.float32, "target")
    pred, target = pred.contiguous(), target.contiguous()
    B, C_ = target.shape[0], target.shape[1]
    HW = target[0, 0].numel()
    loss = torch.empty((), device=pred.device, dtype=torch.float64)
    grad = torch.empty_like(target) if want_grad else None
    check(_lib.lib().tdm_mse_fwd_bwd(ptr(pred), ptr(target), ptr(loss), ptr(grad), float(grad_scale), B, C_,
                                     pred.shape[1], HW, stream_handle()))
    return loss, grad


def sumsq(x, accum, zero_first):
    _req(x, torch.float32, "x")
    check(_lib.lib().tdm_sumsq_f32(ptr(x), x.numel(), ptr(accum), int(zero_first), stream_handle()))


def clip_coef(sumsq_acc, max_norm, inv_scale, coef_out, norm_out=None):
    check(_lib.lib().tdm_clip_coef(ptr(sumsq_acc), float(max_norm), float(inv_scale), ptr(coef_out), ptr(norm_out),
                                   stream_handle()))


def adamw_step(param, grad, exp_avg, exp_avg_sq, param_bf16, lr, beta1, beta2, eps, weight_decay, step, grad_coef=None):
    for t, n in ((param, "param"), (grad, "grad"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
        _req(t, torch.float32, n)
    _req(param_bf16, torch.bfloat16, "param_bf16")
    check(_lib.lib().tdm_adamw_step(ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), ptr(param_bf16), param.numel(),
                                    float(lr), float(beta1), float(beta2), float(eps), float(weight_decay), int(step),
                                    ptr(grad_coef), stream_handle()))
