"""tcgen05 GEMM (tdm_gemm_bf16) against a torch fp32 reference of the same contraction + epilogue.

Floating-point kernel: inputs are bf16, accumulation fp32; tolerance is stated per test."""
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def _mk(shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, device="cuda", generator=g) * scale).to(torch.bfloat16)


def _close(got, ref, tol, what=""):
    err = (got.float() - ref).abs().max().item()
    den = ref.abs().max().item() + 1e-12
    assert err / den <= tol, f"{what}: max abs err {err:.4g} / max ref {den:.4g} = {err/den:.3g} > {tol}"


@pytest.mark.parametrize("cta_pair", [1, 2])
@pytest.mark.parametrize("tile_n", [64, 128, 192, 256])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
def test_gemm_majors(tile_n, a_mn, b_mn, cta_pair):
    from tdm_b200 import ops
    if cta_pair == 2 and (tile_n == 64 or (b_mn and tile_n == 192)):
        pytest.skip("not a CTA-pair configuration")
    M, N, K = 384, 512, 320
    a = _mk((M, K), seed=1)
    b = _mk((N, K), seed=2)
    ref = a.float() @ b.float().t()
    a_s = a.t().contiguous() if a_mn else a
    b_s = b.t().contiguous() if b_mn else b
    out = torch.empty(M, N, device="cuda", dtype=torch.float32)
    ops.gemm(a_s, b_s, a_mn_major=a_mn, b_mn_major=b_mn, out_f32=out, tile_n=tile_n, cta_pair=cta_pair)
    torch.cuda.synchronize()
    _close(out, ref, 1e-5 * K ** 0.5 + 1e-4, f"majors a_mn={a_mn} b_mn={b_mn} bn={tile_n}")


@pytest.mark.parametrize("cta_pair", [0, 1, 2])
@pytest.mark.parametrize("M,N,K", [(100, 72, 40), (8192, 1152, 1152), (960, 2304, 1152), (130, 264, 4096), (512, 32, 256),
                                   (8192, 4608, 1152), (1000, 1152, 200)])
def test_gemm_shapes_and_tails(M, N, K, cta_pair):
    from tdm_b200 import ops
    if cta_pair == 2 and M <= 128:
        pytest.skip("a CTA pair needs more than one 128-row tile")
    a = _mk((M, K), seed=3)
    b = _mk((N, K), 0.05, seed=4)
    ref = a.float() @ b.float().t()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, b, out_bf16=out, cta_pair=cta_pair)
    torch.cuda.synchronize()
    _close(out, ref, 6e-3, f"shape {M}x{N}x{K}")  # bf16 output rounding 2^-8


@pytest.mark.parametrize("cta_pair", [1, 2])
def test_gemm_epilogue_full(cta_pair):
    from tdm_b200 import ops
    B, T, N, K = 3, 200, 256, 192
    M = B * T
    a = _mk((M, K), seed=5)
    w = _mk((N, K), 0.1, seed=6)
    bias = torch.randn(N, device="cuda")
    gate = torch.randn(B, 6 * N, device="cuda")[:, N:2 * N]  # strided view like a chunk of [B,6,D]
    res = torch.randn(M, N, device="cuda")
    pre = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    o32 = torch.empty(M, N, device="cuda")
    o16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, bias=bias, act=ops.ACT_GELU_TANH, pre_out=pre, gate=gate, rows_per_gate=T, res=res,
             out_f32=o32, out_bf16=o16, alpha=0.5, cta_pair=cta_pair)
    torch.cuda.synchronize()
    v = 0.5 * (a.float() @ w.float().t()) + bias
    _close(pre, v, 6e-3, "pre_out")
    ref = torch.nn.functional.gelu(v, approximate="tanh") * gate.repeat_interleave(T, 0) + res
    _close(o32, ref, 2e-4, "out_f32")
    _close(o16, ref, 6e-3, "out_bf16")


def test_gemm_gelu_bwd_and_accumulate():
    from tdm_b200 import ops
    M, N, K = 256, 384, 128
    dy = _mk((M, K), seed=7)
    w = _mk((K, N), 0.1, seed=8)       # stored [N_out=K, K_in=N]: dgrad reads it MN-major
    u = _mk((M, N), seed=9)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(dy, w, b_mn_major=True, act=ops.ACT_GELU_TANH_BWD, act_in=u, out_bf16=out)
    uf = u.float().requires_grad_(True)
    torch.nn.functional.gelu(uf, approximate="tanh").backward(dy.float() @ w.float())
    _close(out, uf.grad, 6e-3, "gelu bwd dgrad")
    # in-place accumulate: res aliases out_f32
    acc = torch.randn(M, N, device="cuda")
    ref = acc + dy.float() @ w.float()
    ops.gemm(dy, w, b_mn_major=True, res=acc, out_f32=acc)
    torch.cuda.synchronize()
    _close(acc, ref, 2e-4, "accumulate")


@pytest.mark.parametrize("cta_pair", [1, 2])
def test_gemm_lora_pair_and_segments(cta_pair):
    from tdm_b200 import ops
    M, D, r = 300, 128, 64
    x = _mk((M, D), seed=10)
    w = _mk((3 * D, D), 0.1, seed=11)
    t = _mk((M, 3 * r), seed=12)            # x A^T for q,k,v
    bl = _mk((3 * D, r), 0.1, seed=13)      # stacked B_q;B_k;B_v
    out = torch.empty(M, 3 * D, device="cuda")
    ops.gemm(x, w, a2=t, b2=bl, seg_n=D, seg_k2=r, out_f32=out, cta_pair=cta_pair)
    torch.cuda.synchronize()
    ref = x.float() @ w.float().t()
    for s in range(3):
        ref[:, s * D:(s + 1) * D] += t[:, s * r:(s + 1) * r].float() @ bl[s * D:(s + 1) * D].float().t()
    _close(out, ref, 2e-4, "segmented LoRA pair")
    # plain (unsegmented) pair with K2 = 16 (< one k-block: TMA zero-fills the rest)
    t2 = _mk((M, 16), seed=14)
    b2 = _mk((D, 16), 0.1, seed=15)
    out2 = torch.empty(M, D, device="cuda")
    ops.gemm(x, w[:D], a2=t2, b2=b2, out_f32=out2, cta_pair=cta_pair)
    torch.cuda.synchronize()
    _close(out2, x.float() @ w[:D].float().t() + t2.float() @ b2.float().t(), 2e-4, "rank-16 pair")


@pytest.mark.parametrize("cta_pair", [1, 2])
@pytest.mark.parametrize("split_k", [0, 3, 7])
def test_gemm_split_k_accumulate(split_k, cta_pair):
    """wgrad-shaped contraction (few output tiles, long K): split-K partial sums are red.add'ed into out_f32."""
    from tdm_b200 import ops, _lib
    tokens, n_out, k_in = 4096, 192, 320
    dy = _mk((tokens, n_out), seed=20)
    x = _mk((tokens, k_in), seed=21)
    acc = torch.randn(n_out, k_in, device="cuda")
    ref = acc + dy.float().t() @ x.float()
    ops.gemm(dy, x, a_mn_major=True, b_mn_major=True, res=acc, out_f32=acc, split_k=split_k, cta_pair=cta_pair, tile_n=128)
    torch.cuda.synchronize()
    _close(acc, ref, 3e-4, f"split_k={split_k}")
    with pytest.raises(_lib.TdmError):  # split-K is only defined for pure accumulation
        ops.gemm(dy, x, a_mn_major=True, b_mn_major=True, out_f32=torch.empty_like(acc), split_k=2)


def test_gemm_all_activations_match_torch():
    from tdm_b200 import ops
    M, N, K = 256, 192, 128
    a, w = _mk((M, K), seed=30), _mk((N, K), 0.2, seed=31)
    u = _mk((M, N), seed=32)
    v = a.float() @ w.float().t()
    F = torch.nn.functional
    uf = u.float()
    sg = torch.sigmoid(uf)
    cases = {ops.ACT_GELU_TANH: F.gelu(v, approximate="tanh"), ops.ACT_SILU: F.silu(v),
             ops.ACT_SILU_BWD: v * (sg * (1 + uf * (1 - sg)))}
    for act, ref in cases.items():
        out = torch.empty(M, N, device="cuda")
        ops.gemm(a, w, act=act, act_in=u if act == ops.ACT_SILU_BWD else None, out_f32=out)
        torch.cuda.synchronize()
        _close(out, ref, 2e-3, f"act {act}")  # tanh.approx / __expf: ~1e-3 relative


def test_gemm_rejects_bad_args():
    from tdm_b200 import ops, _lib
    a = _mk((64, 64))
    b = _mk((60, 64))
    with pytest.raises(_lib.TdmError):
        ops.gemm(a, b, out_f32=torch.empty(64, 60, device="cuda"))  # N % 8 != 0
    with pytest.raises(_lib.TdmError):
        ops.gemm(a, _mk((64, 64)))  # no output


_EPI_CASES = {
    # name: kwargs builder (M, N given) -> the eight feature sets the streaming epilogue is specialised for, plus extras
    "bf16": lambda t: dict(bias=t["bias"], out_bf16=t["o16"]),
    "gelu": lambda t: dict(bias=t["bias"], act=1, out_bf16=t["o16"]),
    "gelu_pre": lambda t: dict(bias=t["bias"], act=1, pre_out=t["pre"], out_bf16=t["o16"]),
    "gelu_bwd": lambda t: dict(act=2, act_in=t["u"], out_bf16=t["o16"]),
    "res32": lambda t: dict(bias=t["bias"], res=t["res"], out_f32=t["o32"]),
    "gate_res32": lambda t: dict(bias=t["bias"], gate=t["gate"], rows_per_gate=t["T"], res=t["res"], out_f32=t["o32"]),
    "gate_res32_bf16": lambda t: dict(bias=t["bias"], gate=t["gate"], rows_per_gate=t["T"], res=t["res"], out_f32=t["o32"],
                                      out_bf16=t["o16"]),
    "merge": lambda t: dict(alpha=0.5, res=t["res"], out_bf16=t["o16"]),
    "everything": lambda t: dict(bias=t["bias"], act=1, pre_out=t["pre"], gate=t["gate"], rows_per_gate=t["T"], res=t["res"],
                                 out_f32=t["o32"], out_bf16=t["o16"], alpha=0.5),
    "silu_bwd_inplace": lambda t: dict(act=4, act_in=t["u"], res=t["o32"], out_f32=t["o32"]),
}


@pytest.mark.parametrize("case", sorted(_EPI_CASES))
@pytest.mark.parametrize("cta_pair,tile_n", [(1, 64), (1, 192), (2, 128), (2, 256)])
def test_streaming_epilogue_instances_are_bit_identical(case, cta_pair, tile_n):
    """The generic epilogue (epilogue=1), the streaming epilogue with run-time tests (2) and its feature-specialised
    instance (3) must agree bit for bit on every output -- ragged M and N tails included -- and match the fp32 formula."""
    from tdm_b200 import ops
    B, T, N, K = 3, 384, 328, 200   # M = 1152 rows (gate uniform per 128-row tile), N tail of 8 columns past 320
    M = B * T
    a, w = _mk((M, K), seed=21), _mk((N, K), 0.1, seed=22)
    base = dict(bias=torch.randn(N, device="cuda"), gate=torch.randn(B, 6 * N, device="cuda")[:, N:2 * N], T=T,
                res=torch.randn(M, N, device="cuda"), u=_mk((M, N), seed=23))
    outs = []
    for mode in (1, 2, 3):
        t = dict(base, pre=torch.zeros(M, N, device="cuda", dtype=torch.bfloat16), o32=base["res"].clone() * 0.25,
                 o16=torch.zeros(M, N, device="cuda", dtype=torch.bfloat16))
        ops.gemm(a, w, tile_n=tile_n, cta_pair=cta_pair, epilogue=mode, **_EPI_CASES[case](t))
        torch.cuda.synchronize()
        outs.append((t["pre"], t["o32"], t["o16"]))
    for mode, o in zip((2, 3), outs[1:]):
        for name, x, y in zip(("pre_out", "out_f32", "out_bf16"), outs[0], o):
            assert torch.equal(x, y), f"{case}: epilogue={mode} differs from the generic epilogue in {name}"
    if case == "gate_res32_bf16":  # one closed-form check of the shared result
        v = a.float() @ w.float().t() + base["bias"]
        ref = v * base["gate"].repeat_interleave(T, 0) + base["res"]
        _close(outs[2][1], ref, 2e-4, "out_f32")
        _close(outs[2][2], ref, 6e-3, "out_bf16")


def test_streaming_epilogue_refuses_what_it_does_not_cover():
    from tdm_b200 import ops, _lib
    a, w = _mk((256, 64), seed=1), _mk((128, 64), seed=2)
    acc = torch.zeros(256, 128, device="cuda")
    with pytest.raises(_lib.TdmError):
        ops.gemm(a, w, res=acc[:100], res_row_mod=100, out_f32=acc, epilogue=3)


def test_gemm_stacked_segments():
    """seg_k: column segments contract over their own window of A / B (LoRA dt for stacked q/k/v); seg_m: row segments read
    their own B columns (LoRA dB).  One launch must equal the per-segment launches bit for bit."""
    from tdm_b200 import ops
    tokens, D, r, nseg = 384, 256, 64, 3
    dy = _mk((tokens, nseg * D), seed=31)
    lb = _mk((nseg * D, r), 0.1, seed=32)          # stacked lora_B: [nseg * D, r]
    t = _mk((tokens, nseg * r), seed=33)
    # dt[:, s] = 0.5 * dy[:, s] @ B_s
    dt_one = torch.zeros(tokens, nseg * r, device="cuda", dtype=torch.bfloat16)
    ops.gemm(dy, lb, b_mn_major=True, alpha=0.5, out_bf16=dt_one, seg_n=r, seg_k=D, tile_n=64, cta_pair=1)
    dt_ref = torch.zeros_like(dt_one)
    for s in range(nseg):
        ops.gemm(dy[:, s * D:(s + 1) * D], lb[s * D:(s + 1) * D], b_mn_major=True, alpha=0.5, out_bf16=dt_ref[:, s * r:(s + 1) * r])
    torch.cuda.synchronize()
    assert torch.equal(dt_one, dt_ref)
    _close(dt_one[:, r:2 * r], 0.5 * dy[:, D:2 * D].float() @ lb[D:2 * D].float(), 6e-3, "dt segment 1")
    # dB_s += dy[:, s]^T t[:, s]
    g_one = torch.ones(nseg * D, r, device="cuda")
    ops.gemm(dy, t, a_mn_major=True, b_mn_major=True, res=g_one, out_f32=g_one, seg_m=D, seg_nb=r, tile_n=64, cta_pair=1, split_k=1)
    g_ref = torch.ones(nseg * D, r, device="cuda")
    for s in range(nseg):
        ops.gemm(dy[:, s * D:(s + 1) * D], t[:, s * r:(s + 1) * r], a_mn_major=True, b_mn_major=True, res=g_ref[s * D:(s + 1) * D],
                 out_f32=g_ref[s * D:(s + 1) * D], tile_n=64, cta_pair=1, split_k=1)
    torch.cuda.synchronize()
    assert torch.equal(g_one, g_ref)
    _close(g_one[D:2 * D], 1.0 + dy[:, D:2 * D].float().t() @ t[:, r:2 * r].float(), 2e-4, "dB segment 1")


def test_lora_merge_kernel_against_torch():
    """tdm_lora_merge on a table of three problems of different shapes: W_eff = bf16(W32 + s * B A) with bf16 adapters,
    fp32 accumulation -- within one bf16 ulp of the fp32 torch result, every tile of every problem written."""
    import struct
    from tdm_b200 import _lib
    from tdm_b200._lib import check, ptr, stream_handle
    torch.manual_seed(3)
    dev = "cuda"
    shapes = [(256, 384), (1152, 1152), (128, 4608), (4608, 1152)]
    scale = 0.75
    rows, tile, keep = [], 0, []
    for n_out, k_in in shapes:
        w32 = torch.randn(n_out, k_in, device=dev)
        la = (torch.randn(64, k_in, device=dev) * 0.2).bfloat16()
        lb = (torch.randn(n_out, 64, device=dev) * 0.2).bfloat16()
        weff = torch.full((n_out, k_in), float("nan"), device=dev, dtype=torch.bfloat16)
        keep.append((w32, la, lb, weff))
        rows.append(struct.pack("<QQQQiiii", w32.data_ptr(), weff.data_ptr(), la.data_ptr(), lb.data_ptr(), n_out, k_in, tile, 0))
        tile += (n_out // 128) * (k_in // 128)
    table = torch.frombuffer(bytearray(b"".join(rows)), dtype=torch.uint8).to(dev)
    for _ in range(3):   # relaunches reuse the ring / staging buffers of a persistent CTA
        check(_lib.lib().tdm_lora_merge(ptr(table), len(rows), tile, scale, stream_handle()))
    torch.cuda.synchronize()
    for w32, la, lb, weff in keep:
        ref = w32 + scale * (lb.float() @ la.float())
        assert torch.isfinite(weff.float()).all()
        err = (weff.float() - ref).abs()
        assert (err <= ref.abs() * 2.0 ** -8 + 1e-6).all(), err.max().item()


@pytest.mark.parametrize("M,N,K,b_mn", [(8192, 1152, 256, False), (8192, 1152, 256, True), (8192, 3456, 128, False),
                                        (4096, 640, 192, True), (2304, 1152, 320, False), (19200, 384, 128, False)])
def test_gemm_pair_narrow_tail_and_balanced_order(M, N, K, b_mn, monkeypatch):
    """CTA-pair kernel, 256-column tiles, last column tile <= 128 columns wide: that tile runs N = 128 MMAs and the unit
    order deals the narrow tiles to the pairs with one unit more (gemm_plan_balance).  Same values as the identity order
    with full-width MMAs (TDM_B200_GEMM_BALANCE=0), bit for bit, with a residual + gate + two-output epilogue, and within
    bf16 tolerance of the fp32 product."""
    from tdm_b200 import ops
    a = _mk((M, K), seed=31)
    w = _mk((K, N) if b_mn else (N, K), 0.1, seed=32)
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    rows_per_gate = 128
    gate = torch.randn((M + rows_per_gate - 1) // rows_per_gate, N, device="cuda")
    outs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("TDM_B200_GEMM_BALANCE", flag)
        o32 = torch.full((M, N), float("nan"), device="cuda")
        o16 = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
        ops.gemm(a, w, b_mn_major=b_mn, bias=bias, gate=gate, rows_per_gate=rows_per_gate, res=res, out_f32=o32, out_bf16=o16,
                 tile_n=256, cta_pair=2)
        torch.cuda.synchronize()
        outs.append((o32, o16))
    wt = w.float() if b_mn else w.float().t()
    ref = (a.float() @ wt + bias) * gate.repeat_interleave(rows_per_gate, 0)[:M] + res
    assert torch.isfinite(outs[0][0]).all()
    assert (outs[0][0] - ref).abs().max().item() < 2e-2 * ref.abs().max().item()
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
