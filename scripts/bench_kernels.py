"""Kernel micro-benchmarks on the PixArt-alpha-512 shapes at batch 8 (CUDA events on the launching stream,
rotating buffers larger than L2).  One JSON line per case; `tflops` uses algorithmic FLOPs, `gbs` algorithmic bytes."""
import json
import sys
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
from tdm_b200 import ops

DEV = "cuda"
NB = 4


from _timing import timeit as _timeit


def timeit(fn, iters=20, warm=3):
    return _timeit(fn, iters, warm)[0]


def gemm_case(name, M, N, K, epi, tile_n=0, cta_pair=0):
    B, T = 8, M // 8
    a = [torch.randn(M, K, device=DEV).bfloat16() for _ in range(NB)]
    w = [(torch.randn(N, K, device=DEV) * 0.05).bfloat16() for _ in range(NB)]
    bias = torch.randn(N, device=DEV)
    gate = torch.randn(B, 6 * N, device=DEV)[:, N:2 * N]
    res = [torch.randn(M, N, device=DEV) for _ in range(NB)]
    o32 = [torch.empty(M, N, device=DEV) for _ in range(NB)]
    o16 = [torch.empty(M, N, device=DEV, dtype=torch.bfloat16) for _ in range(NB)]
    pre = [torch.empty(M, N, device=DEV, dtype=torch.bfloat16) for _ in range(NB)]
    kw = {
        "plain": lambda i: dict(out_bf16=o16[i]),
        "bias": lambda i: dict(bias=bias, out_bf16=o16[i]),
        "gelu": lambda i: dict(bias=bias, act=ops.ACT_GELU_TANH, out_bf16=o16[i]),
        "gelu_save": lambda i: dict(bias=bias, act=ops.ACT_GELU_TANH, pre_out=pre[i], out_bf16=o16[i]),
        "gelu_bwd": lambda i: dict(act=ops.ACT_GELU_TANH_BWD, act_in=pre[i], out_bf16=o16[i]),
        "gate_res": lambda i: dict(bias=bias, gate=gate, rows_per_gate=T, res=res[i], out_f32=o32[i]),
        "gate_res_2out": lambda i: dict(bias=bias, gate=gate, rows_per_gate=T, res=res[i], out_f32=o32[i], out_bf16=o16[i]),
        "res": lambda i: dict(bias=bias, res=res[i], out_f32=o32[i]),
        "res_inplace": lambda i: dict(res=o32[i], out_f32=o32[i]),
    }[epi]
    us = timeit(lambda i: ops.gemm(a[i % NB], w[i % NB], tile_n=tile_n, cta_pair=cta_pair, **kw(i % NB)))
    print(json.dumps({"kernel": "gemm", "case": name, "epi": epi, "M": M, "N": N, "K": K, "tile_n": tile_n, "cta_pair": cta_pair, "us": round(us, 2),
                      "tflops": round(2.0 * M * N * K / us / 1e6, 1)}), flush=True)


def wgrad_case(name, tokens, n_out, k_in, split_k=0):
    dy = [torch.randn(tokens, n_out, device=DEV).bfloat16() for _ in range(NB)]
    x = [torch.randn(tokens, k_in, device=DEV).bfloat16() for _ in range(NB)]
    dw = torch.zeros(n_out, k_in, device=DEV)
    us = timeit(lambda i: ops.gemm(dy[i % NB], x[i % NB], a_mn_major=True, b_mn_major=True, res=dw, out_f32=dw, split_k=split_k))
    print(json.dumps({"kernel": "gemm_wgrad", "case": name, "tokens": tokens, "n_out": n_out, "k_in": k_in, "split_k": split_k,
                      "us": round(us, 2), "tflops": round(2.0 * tokens * n_out * k_in / us / 1e6, 1)}), flush=True)


def attn_case(name, B, H, Sq, Sk, d, bwd=False, bias=False):
    D = H * d
    qkv = [torch.randn(B, Sq, 3 * D, device=DEV).bfloat16() for _ in range(NB)] if Sq == Sk else None
    if qkv is not None:
        q = [t[:, :, :D].view(B, Sq, H, d) for t in qkv]
        k = [t[:, :, D:2 * D].view(B, Sk, H, d) for t in qkv]
        v = [t[:, :, 2 * D:].view(B, Sk, H, d) for t in qkv]
    else:
        q = [torch.randn(B, Sq, H, d, device=DEV).bfloat16() for _ in range(NB)]
        kv = [torch.randn(B, Sk, 2 * D, device=DEV).bfloat16() for _ in range(NB)]
        k = [t[:, :, :D].view(B, Sk, H, d) for t in kv]
        v = [t[:, :, D:].view(B, Sk, H, d) for t in kv]
    o = [torch.empty(B, Sq, H, d, device=DEV, dtype=torch.bfloat16) for _ in range(NB)]
    lse = [torch.empty(B, H, Sq, device=DEV) for _ in range(NB)]
    bs = torch.zeros(B, Sk, device=DEV) if bias else None
    scale = d ** -0.5
    fl = 4.0 * B * H * Sq * Sk * d
    if not bwd:
        us = timeit(lambda i: ops.attn_fwd(q[i % NB], k[i % NB], v[i % NB], o[i % NB], scale, bias=bs, lse=lse[i % NB]))
    else:
        for i in range(NB):
            ops.attn_fwd(q[i], k[i], v[i], o[i], scale, bias=bs, lse=lse[i])
        do = [torch.randn_like(t) for t in o]
        dq = [torch.empty(B, Sq, H, d, device=DEV, dtype=torch.bfloat16) for _ in range(NB)]
        dk = [torch.empty(B, Sk, H, d, device=DEV, dtype=torch.bfloat16) for _ in range(NB)]
        dv = [torch.empty(B, Sk, H, d, device=DEV, dtype=torch.bfloat16) for _ in range(NB)]
        us = timeit(lambda i: ops.attn_bwd(q[i % NB], k[i % NB], v[i % NB], o[i % NB], do[i % NB], lse[i % NB], scale,
                                           dq[i % NB], dk[i % NB], dv[i % NB], bias=bs))
        fl *= 2.5
    print(json.dumps({"kernel": "attn_bwd" if bwd else "attn_fwd", "case": name, "B": B, "H": H, "Sq": Sq, "Sk": Sk, "d": d,
                      "us": round(us, 2), "tflops": round(fl / us / 1e6, 1)}), flush=True)


def ln_case(M, D):
    x = [torch.randn(M, D, device=DEV) for _ in range(NB)]
    mod = torch.randn(8, 6 * D, device=DEV)
    us = timeit(lambda i: ops.ln_modulate_fwd(x[i % NB], mod[:, :D], mod[:, D:2 * D], M // 8))
    print(json.dumps({"kernel": "ln_modulate_fwd", "M": M, "D": D, "us": round(us, 2), "gbs": round(M * D * 6 / us / 1e3, 1)}), flush=True)
    dh = [torch.randn(M, D, device=DEV).bfloat16() for _ in range(NB)]
    _, mean, rstd = ops.ln_modulate_fwd(x[0], mod[:, :D], mod[:, D:2 * D], M // 8)
    dx_in = torch.randn(M, D, device=DEV)
    us = timeit(lambda i: ops.ln_modulate_bwd(dh[i % NB], x[i % NB], mean, rstd, mod[:, D:2 * D], M // 8, dx_in=dx_in, want_bf16=True))
    print(json.dumps({"kernel": "ln_modulate_bwd", "M": M, "D": D, "us": round(us, 2), "gbs": round(M * D * 16 / us / 1e3, 1)}), flush=True)


def solver_case(B):
    """Solver step / re-noise / loss at the real batch and at a large synthetic batch (HBM bandwidth visible)."""
    from tdm_b200.scheduler import DDPMScheduler
    sch = DDPMScheduler(beta_start=1e-4, beta_end=0.02, beta_schedule="linear", steps_offset=1, clip_sample=False)
    a = torch.sqrt(sch.alphas_cumprod).to(DEV).bfloat16()
    sg = torch.sqrt(1 - sch.alphas_cumprod).to(DEV).bfloat16()
    acp = sch.alphas_cumprod.to(DEV).float().contiguous()
    out = [torch.randn(B, 8, 64, 64, device=DEV) for _ in range(NB)]
    x = [torch.randn(B, 4, 64, 64, device=DEV).bfloat16() for _ in range(NB)]
    t = torch.randint(1, 1000, (B,), device=DEV)
    n = B * 4 * 64 * 64
    us = timeit(lambda i: ops.solver_step(out[i % NB], x[i % NB], t, a, sg, "epsilon", t_next=t - 1, alphas_cumprod=acp,
                                          next_dtype=torch.bfloat16, want_eps=False))
    print(json.dumps({"kernel": "solver_step", "B": B, "us": round(us, 2), "gbs": round(n * (4 + 2 + 4 + 2) / us / 1e3, 1)}), flush=True)
    nz = [torch.randn(B, 4, 64, 64, device=DEV).bfloat16() for _ in range(NB)]
    t2 = torch.clamp(t + 50, max=999)
    us = timeit(lambda i: ops.renoise(x[i % NB], nz[i % NB], t, t2, a, sg, torch.bfloat16))
    print(json.dumps({"kernel": "renoise", "B": B, "us": round(us, 2), "gbs": round(n * 6 / us / 1e3, 1)}), flush=True)
    m = [torch.randn(B, 4, 64, 64, device=DEV) for _ in range(3)]
    us = timeit(lambda i: ops.tdm_loss(m[0], m[1], m[2], 4.5, True))
    print(json.dumps({"kernel": "tdm_loss_fwd_bwd", "B": B, "us": round(us, 2), "gbs": round(n * (3 * 4 * 2 + 4) / us / 1e3, 1)}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "wgrad", "attn", "ln", "solver"]
    M, D = 8192, 1152
    if "pair" in which:
        for nm, (m, n, k), epi in (("qkv", (M, 3 * D, D), "bias"), ("q2", (M, D, D), "bias"), ("o_proj", (M, D, D), "gate_res_2out"),
                                   ("ff1", (M, 4 * D, D), "gelu"), ("ff1_plain", (M, 4 * D, D), "plain"), ("ff2", (M, D, 4 * D), "gate_res"),
                                   ("big", (8192, 8192, 8192), "plain")):
            for pr in (1, 2):
                for tn in (128, 192, 256):
                    gemm_case(nm, m, n, k, epi, tn, pr)
    if "gemm" in which:
        gemm_case("qkv", M, 3 * D, D, "bias")
        gemm_case("qkv_192", M, 3 * D, D, "bias", 192)
        gemm_case("qkv_256", M, 3 * D, D, "bias", 256)
        gemm_case("o_proj", M, D, D, "gate_res_2out")
        gemm_case("o_proj_192", M, D, D, "gate_res_2out", 192)
        gemm_case("o_proj_256", M, D, D, "gate_res_2out", 256)
        gemm_case("q2", M, D, D, "bias")
        gemm_case("o2", M, D, D, "res")
        gemm_case("ff1", M, 4 * D, D, "gelu")
        gemm_case("ff1_save", M, 4 * D, D, "gelu_save")
        gemm_case("ff1_plain", M, 4 * D, D, "plain")
        gemm_case("ff2", M, D, 4 * D, "gate_res")
        gemm_case("ff2_plain", M, D, 4 * D, "plain")
        gemm_case("ff2_dgrad", M, 4 * D, D, "gelu_bwd")
        gemm_case("kv2", 960, 2 * D, D, "bias")
        gemm_case("lora_t", M, 64, D, "plain")
        gemm_case("lora_t_qkv", M, 192, D, "plain")
    if "wgrad" in which:
        wgrad_case("w_o", M, D, D)
        wgrad_case("w_o_nosplit", M, D, D, 1)
        wgrad_case("w_o_split2", M, D, D, 2)
        wgrad_case("w_o_split3", M, D, D, 3)
        wgrad_case("w_qkv_split2", M, 3 * D, D, 2)
        wgrad_case("w_qkv", M, 3 * D, D)
        wgrad_case("w_ff1", M, 4 * D, D)
        wgrad_case("w_ff2", M, D, 4 * D)
        wgrad_case("lora_dB", M, D, 64)
        wgrad_case("lora_dA", M, 64, D)
    if "attn" in which:
        attn_case("self", 8, 16, 1024, 1024, 72)
        attn_case("cross", 8, 16, 1024, 120, 72, bias=True)
        attn_case("self_1024px", 2, 16, 4096, 4096, 72)
        attn_case("self_d64", 8, 16, 1024, 1024, 64)
        attn_case("self", 8, 16, 1024, 1024, 72, bwd=True)
        attn_case("cross", 8, 16, 1024, 120, 72, bwd=True, bias=True)
    if "ln" in which:
        ln_case(M, D)
    if "solver" in which:
        solver_case(8)
        solver_case(64)
        solver_case(2048)
