"""HBM-bound block kernels and tcgen05 attention through the C ABI against fp32 torch references of the
same ops (floating-point kernels: tolerances stated per test; bf16 outputs carry 2^-8 rounding)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def _rel(got, ref):
    return ((got.float() - ref.float()).abs().max() / (ref.float().abs().max() + 1e-12)).item()


@pytest.mark.parametrize("B,N,D", [(2, 256, 256), (8, 1024, 1152), (3, 100, 1152)])
def test_ln_modulate_fwd_bwd(B, N, D):
    from tdm_b200 import ops
    torch.manual_seed(0)
    M = B * N
    x = (torch.randn(M, D, device="cuda") * 2 + 0.3).requires_grad_(True)
    mod = (torch.randn(B, 6, D, device="cuda") * 0.3).requires_grad_(True)
    shift, scale = mod[:, 0], mod[:, 1]
    h, mean, rstd = ops.ln_modulate_fwd(x.detach(), shift.detach(), scale.detach(), N, 1e-6)
    ref = F.layer_norm(x.view(B, N, D), (D,), eps=1e-6) * (1 + scale[:, None]) + shift[:, None]
    assert _rel(h.view(B, N, D), ref) < 6e-3
    # backward
    dh = torch.randn(M, D, device="cuda").to(torch.bfloat16)
    dx_in = torch.randn(M, D, device="cuda")
    ref.backward(dh.float().view(B, N, D))
    dmod = torch.zeros(B, 6, D, device="cuda")
    dx, dxb = ops.ln_modulate_bwd(dh, x.detach(), mean, rstd, scale.detach(), N, dx_in=dx_in, want_bf16=True,
                                  dshift=dmod[:, 0], dscale=dmod[:, 1])
    assert _rel(dx - dx_in, x.grad) < 2e-5
    assert _rel(dxb, x.grad + dx_in) < 6e-3
    assert _rel(dmod[:, 0], mod.grad[:, 0]) < 1e-4
    assert _rel(dmod[:, 1], mod.grad[:, 1]) < 1e-4


def test_gate_bwd_and_colsum():
    from tdm_b200 import ops
    torch.manual_seed(1)
    B, N, D = 4, 256, 384
    M = B * N
    dx = torch.randn(M, D, device="cuda")
    y = torch.randn(M, D, device="cuda").to(torch.bfloat16)
    mod = torch.randn(B, 6, D, device="cuda")
    gate = mod[:, 2]
    dgate = torch.zeros(B, 6, D, device="cuda")
    dy = ops.gate_bwd(dx, gate, N, y=y, dgate=dgate[:, 2])
    assert _rel(dy.view(B, N, D), dx.view(B, N, D) * gate[:, None]) < 6e-3
    ref = (dx.view(B, N, D) * y.float().view(B, N, D)).sum(1)
    assert _rel(dgate[:, 2], ref) < 1e-4
    assert dgate[:, 0].abs().max().item() == 0
    dy2 = ops.gate_bwd(dx, gate, N)  # frozen gate: no reduction
    assert torch.equal(dy2, dy)
    out = torch.zeros(D, device="cuda")
    ops.colsum(y, out)
    assert _rel(out, y.float().sum(0)) < 1e-4


def test_modulation_and_sinusoid_and_patches():
    from tdm_b200 import ops
    from oracle.dit import timestep_sinusoid
    torch.manual_seed(2)
    B, D = 3, 256
    table = torch.randn(6, D, device="cuda")
    tv = torch.randn(B, 6 * D, device="cuda")
    out = ops.modulation(table, tv, True)
    assert torch.equal(out, table[None] + tv.view(B, 6, D))
    table2 = torch.randn(2, D, device="cuda")
    emb = torch.randn(B, D, device="cuda")
    out2 = ops.modulation(table2, emb, False)
    assert torch.equal(out2, table2[None] + emb[:, None])
    # backward of both forms
    dmod = torch.randn(B, 6, D, device="cuda")
    dtable = torch.zeros(6, D, device="cuda")
    dtv = torch.zeros(B, 6 * D, device="cuda")
    ops.modulation_bwd(dmod, dtable, dtv, True, False)
    assert _rel(dtable, dmod.sum(0)) < 1e-6 and torch.equal(dtv.view(B, 6, D), dmod)
    dm2 = torch.randn(B, 2, D, device="cuda")
    dt2 = torch.zeros(2, D, device="cuda")
    demb = torch.ones(B, D, device="cuda")
    ops.modulation_bwd(dm2, dt2, demb, False, True)
    assert _rel(demb, 1 + dm2.sum(1)) < 1e-6
    # sinusoid (bf16 output of an fp32 formula; |x| <= 1 so abs tol = bf16 half-ulp + sincos range-reduction error)
    t = torch.tensor([0, 1, 224, 899, 999], device="cuda")
    s = ops.timestep_sinusoid(t)
    assert (s.float() - timestep_sinusoid(t)).abs().max().item() < 5e-3
    # patchify == unfold of the conv input; unpatchify == the einsum of Appendix A step 6
    lat = torch.randn(2, 4, 8, 8, device="cuda").to(torch.float16)
    pt = ops.patchify(lat, 2)
    ref = F.unfold(lat.float(), kernel_size=2, stride=2).transpose(1, 2).reshape(-1, 16)
    assert torch.equal(pt.float(), ref.to(torch.bfloat16).float())
    y = torch.randn(2 * 16, 32, device="cuda")
    up = ops.unpatchify(y, 2, 8, 8, 8, 2)
    ref = torch.einsum("nhwpqc->nchpwq", y.view(2, 4, 4, 2, 2, 8)).reshape(2, 8, 8, 8)
    assert torch.equal(up, ref)
    dout = torch.randn(2, 4, 8, 8, device="cuda")
    dy = ops.unpatchify_bwd(dout, 8, 2)
    yy = y.clone().requires_grad_(True)
    torch.einsum("nhwpqc->nchpwq", yy.view(2, 4, 4, 2, 2, 8)).reshape(2, 8, 8, 8)[:, :4].backward(dout)
    assert torch.equal(dy.float(), yy.grad.to(torch.bfloat16).float())
    xf = torch.randn(1000, device="cuda")
    assert torch.equal(ops.cast_bf16(xf), xf.to(torch.bfloat16))


def _attn_ref(q, k, v, scale, bias):
    # q,k,v: [B,S,H,d] fp32
    s = torch.einsum("bqhd,bkhd->bhqk", q, k) * scale
    if bias is not None:
        s = s + bias[:, None, None, :]
    p = torch.softmax(s, dim=-1)
    return torch.einsum("bhqk,bkhd->bqhd", p, v), torch.logsumexp(s, dim=-1)


@pytest.mark.parametrize("B,H,Sq,Sk,d,cross", [
    (2, 4, 256, 256, 64, False),     # tiny config
    (2, 16, 1024, 1024, 72, False),  # PixArt-512 self-attention
    (2, 16, 1024, 120, 72, True),    # PixArt cross-attention to 120 T5 tokens with a key mask
    (1, 2, 200, 77, 72, True),       # ragged sizes
    (1, 2, 384, 300, 40, True),
    (1, 2, 256, 256, 128, False),    # head_dim > 80: the wide backward kernel
    (1, 3, 130, 250, 96, True),
    (2, 16, 4096, 4096, 72, False),  # PixArt-1024 sequence length (BASELINE.json configs[4])
])
def test_attention_fwd_bwd(B, H, Sq, Sk, d, cross):
    from tdm_b200 import ops
    torch.manual_seed(4)
    D = H * d
    scale = d ** -0.5
    if not cross:
        qkv = torch.randn(B * Sq, 3 * D, device="cuda").to(torch.bfloat16)
        q = qkv[:, :D].view(B, Sq, H, d)
        k = qkv[:, D:2 * D].view(B, Sk, H, d)
        v = qkv[:, 2 * D:].view(B, Sk, H, d)
        bias = None
    else:
        qm = torch.randn(B * Sq, D, device="cuda").to(torch.bfloat16)
        kvm = torch.randn(B * Sk, 2 * D, device="cuda").to(torch.bfloat16)
        q = qm.view(B, Sq, H, d)
        k = kvm[:, :D].view(B, Sk, H, d)
        v = kvm[:, D:].view(B, Sk, H, d)
        lens = torch.randint(1, Sk + 1, (B,), device="cuda")
        mask = (torch.arange(Sk, device="cuda")[None] < lens[:, None]).float()
        bias = (1 - mask) * -10000.0
    o = torch.empty(B, Sq, H, d, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, H, Sq, device="cuda")
    ops.attn_fwd(q, k, v, o, scale, bias=bias, lse=lse)
    torch.cuda.synchronize()
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    o_ref, lse_ref = _attn_ref(qf, kf, vf, scale, bias)
    assert _rel(o, o_ref) < 1e-2, "attention forward"
    assert (lse - lse_ref).abs().max().item() < 2e-3
    # backward
    do = torch.randn(B, Sq, H, d, device="cuda").to(torch.bfloat16)
    o_ref.backward(do.float())
    if not cross:
        dqkv = torch.zeros(B * Sq, 3 * D, device="cuda", dtype=torch.bfloat16)
        dq = dqkv[:, :D].view(B, Sq, H, d)
        dk = dqkv[:, D:2 * D].view(B, Sk, H, d)
        dv = dqkv[:, 2 * D:].view(B, Sk, H, d)
    else:
        dq = torch.zeros(B, Sq, H, d, device="cuda", dtype=torch.bfloat16)
        dkv = torch.zeros(B * Sk, 2 * D, device="cuda", dtype=torch.bfloat16)
        dk = dkv[:, :D].view(B, Sk, H, d)
        dv = dkv[:, D:].view(B, Sk, H, d)
    ops.attn_bwd(q, k, v, o, do, lse, scale, dq, dk, dv, bias=bias)
    torch.cuda.synchronize()
    assert _rel(dv, vf.grad) < 2e-2, "dV"
    assert _rel(dk, kf.grad) < 2e-2, "dK"
    assert _rel(dq, qf.grad) < 2e-2, "dQ"


@pytest.mark.parametrize("B,H,Sq,Sk,d,cross,big", [
    (2, 16, 1024, 1024, 72, False, False),
    (2, 16, 1024, 120, 72, True, False),
    (1, 2, 200, 77, 72, True, False),     # ragged: second query tile of the pair partly / fully out of range
    (1, 3, 130, 250, 96, True, False),
    (1, 2, 384, 300, 40, True, False),
    (3, 5, 640, 384, 64, False, False),   # odd item count: CTAs with different numbers of items
    (2, 4, 512, 1024, 72, False, True),   # row maxima that grow by far more than 2^8 between tiles: the lazy-rescale path
])
def test_attention_fwd_two_stream_kernel_matches_one_tile_kernel(B, H, Sq, Sk, d, cross, big, monkeypatch):
    """attention_fwd2.cu (two query tiles per CTA, default for head_dim <= 96) against attention_fwd.cu
    (TDM_B200_ATTN_V1=1) and the fp32 reference: same softmax arithmetic (same maxima, same exponentials, same bf16 P),
    so the outputs agree to the last bf16 bit or very nearly; both are within the forward tolerance of the reference."""
    from tdm_b200 import ops
    torch.manual_seed(5)
    q = torch.randn(B, Sq, H, d, device="cuda").bfloat16()
    k = torch.randn(B, Sk, H, d, device="cuda").bfloat16()
    v = torch.randn(B, Sk, H, d, device="cuda").bfloat16()
    if big:
        k[:, Sk // 2:] *= 6.0
    bias = None
    if cross:
        lens = torch.randint(1, Sk + 1, (B,), device="cuda")
        bias = (1 - (torch.arange(Sk, device="cuda")[None] < lens[:, None]).float()) * -10000.0
    outs = {}
    for v1 in ("1", "0"):
        monkeypatch.setenv("TDM_B200_ATTN_V1", v1)
        o = torch.full((B, Sq, H, d), float("nan"), device="cuda", dtype=torch.bfloat16)
        lse = torch.full((B, H, Sq), float("nan"), device="cuda")
        ops.attn_fwd(q, k, v, o, d ** -0.5, bias=bias, lse=lse)
        torch.cuda.synchronize()
        outs[v1] = (o, lse)
    o_ref, lse_ref = _attn_ref(q.float(), k.float(), v.float(), d ** -0.5, bias)
    for o, lse in outs.values():
        assert _rel(o, o_ref) < 1e-2
        assert (lse - lse_ref).abs().max().item() < 2e-3
    assert (outs["0"][0].float() - outs["1"][0].float()).abs().max().item() <= 2e-2 * o_ref.abs().max().item() / 4
    assert (outs["0"][1] - outs["1"][1]).abs().max().item() < 1e-5


@pytest.mark.parametrize("B,H,S,d", [(2, 16, 1024, 72), (8, 16, 1024, 72)])
def test_attention_fwd_is_deterministic_with_cold_loads(B, H, S, d):
    """Regression for the P V completion barrier of attn_fwd_kernel: with one barrier for every tile, a parity wait could
    pass two phases early when the V tiles arrived late (cold L2 / TLB, programmatic dependent launch) and O was read
    before the last P V landed -- O changed from run to run while the log-sum-exp stayed the same.  Many launches,
    L2 flushed in between, several items per CTA: every output must equal the first one bit for bit."""
    from tdm_b200 import ops
    D = H * d
    g = torch.Generator(device="cuda").manual_seed(11)
    qkv = torch.randn(B, S, 3 * D, device="cuda", generator=g).bfloat16()
    q, k, v = (qkv[:, :, i * D:(i + 1) * D].view(B, S, H, d) for i in range(3))
    junk = torch.empty(160 << 20, device="cuda", dtype=torch.uint8)   # larger than the 126 MB L2
    first_o = first_lse = None
    for it in range(24):
        o = torch.empty(B, S, H, d, device="cuda", dtype=torch.bfloat16)
        lse = torch.empty(B, H, S, device="cuda")
        junk.zero_()
        ops.attn_fwd(q, k, v, o, d ** -0.5, lse=lse)
        ops.attn_fwd(q, k, v, o, d ** -0.5, lse=lse)   # back to back: the second launch overlaps the first one's tail
        torch.cuda.synchronize()
        if first_o is None:
            first_o, first_lse = o, lse
            ref = F.scaled_dot_product_attention(q.transpose(1, 2).float(), k.transpose(1, 2).float(),
                                                 v.transpose(1, 2).float()).transpose(1, 2)
            assert (o.float() - ref).abs().max().item() < 1e-2
        else:
            assert torch.equal(o, first_o), f"launch {it}: attention output changed"
            assert torch.equal(lse, first_lse)


def test_attention_fwd_bwd_two_streams_many_launches():
    """Barrier-protocol regression (a protocol bug traps after a bounded wait and surfaces as a launch failure): forward +
    backward, self + cross, on two streams at once, L2 flushed now and then.  Two hazards of this kind were found by a
    benchmark run, not by the single-launch tests: a barrier whose producer can run two phases ahead of a waiter needs one
    instance per phase parity (attn_bwd dpt_full, attn_fwd2 o_staged)."""
    from tdm_b200 import ops
    torch.manual_seed(9)
    B, H, S, L, d = 4, 16, 1024, 120, 72
    D = H * d

    def make(cross):
        if not cross:
            qkv = torch.randn(B, S, 3 * D, device="cuda").bfloat16()
            q, k, v = (qkv[:, :, i * D:(i + 1) * D].view(B, S, H, d) for i in range(3))
            dqkv = torch.zeros(B, S, 3 * D, device="cuda", dtype=torch.bfloat16)
            dq, dk, dv = (dqkv[:, :, i * D:(i + 1) * D].view(B, S, H, d) for i in range(3))
            bias = None
        else:
            q = torch.randn(B, S, H, d, device="cuda").bfloat16()
            kv = torch.randn(B, L, 2 * D, device="cuda").bfloat16()
            k, v = kv[:, :, :D].view(B, L, H, d), kv[:, :, D:].view(B, L, H, d)
            dq = torch.zeros(B, S, H, d, device="cuda", dtype=torch.bfloat16)
            dkv = torch.zeros(B, L, 2 * D, device="cuda", dtype=torch.bfloat16)
            dk, dv = dkv[:, :, :D].view(B, L, H, d), dkv[:, :, D:].view(B, L, H, d)
            bias = torch.zeros(B, L, device="cuda")
        o = torch.empty(B, S, H, d, device="cuda", dtype=torch.bfloat16)
        return q, k, v, o, torch.empty(B, H, S, device="cuda"), torch.randn_like(o), dq, dk, dv, bias

    sets = [[make(False), make(True)] for _ in range(2)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    junk = torch.empty(160 << 20, device="cuda", dtype=torch.uint8)
    torch.cuda.synchronize()
    first = None
    for it in range(150):
        for st, cases in zip(streams, sets):
            with torch.cuda.stream(st):
                for (q, k, v, o, lse, do, dq, dk, dv, bias) in cases:
                    ops.attn_fwd(q, k, v, o, d ** -0.5, bias=bias, lse=lse)
                    ops.attn_bwd(q, k, v, o, do, lse, d ** -0.5, dq, dk, dv, bias=bias)
        if it % 5 == 0:
            junk.zero_()
        if it == 0:
            torch.cuda.synchronize()
            first = [c[7].clone() for c in sets[0]]   # dK of both cases: no atomics on that path
    torch.cuda.synchronize()
    for c, f in zip(sets[0], first):
        assert torch.equal(c[7], f), "dK changed between identical launches"
