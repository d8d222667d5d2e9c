"""Oracle DiT (oracle/dit.py, a restatement of the un-vendored diffusers PixArt transformer) checked
against independent implementations available in this image and against the
known answers of SURVEY.md 8c.  CPU only."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import dit
from oracle.dit import DiTConfig, OracleDiT, TINY, PIXART_XL_2_512, count_params
from oracle.schedule import OracleDDPMScheduler


def test_known_answers():
    assert count_params(PIXART_XL_2_512) == 610_856_096
    s = OracleDDPMScheduler()
    a, sg = torch.sqrt(s.alphas_cumprod), torch.sqrt(1 - s.alphas_cumprod)
    assert abs(a[899].item() - 0.016589) < 2e-6 and abs(sg[899].item() - 0.999862) < 2e-6
    assert abs(a[999].item() - 0.006353) < 2e-6
    tiny = OracleDiT(TINY)
    assert sum(p.numel() for p in tiny.parameters()) == count_params(TINY)


def test_sincos_matches_vit_mae():
    mae = pytest.importorskip("transformers.models.vit_mae.modeling_vit_mae")
    ours = dit.sincos_2d(64, 8, base_size=8, interpolation_scale=1)
    ref = mae.get_2d_sincos_pos_embed(64, 8, add_cls_token=False)
    ref = ref.numpy() if hasattr(ref, "numpy") else np.asarray(ref)
    assert np.allclose(ours, ref, atol=1e-6)


def test_timestep_sinusoid_definition():
    t = torch.tensor([0, 1, 500, 999])
    e = dit.timestep_sinusoid(t, 256)
    freq = torch.exp(-math.log(10000.0) * torch.arange(128, dtype=torch.float32) / 128)
    assert torch.allclose(e[:, :128], torch.cos(t.float()[:, None] * freq), atol=1e-6)   # flip_sin_to_cos
    assert torch.allclose(e[:, 128:], torch.sin(t.float()[:, None] * freq), atol=1e-6)


def test_attention_matches_sdpa():
    torch.manual_seed(0)
    att = dit.Attention(128, 4)
    x, ctx = torch.randn(2, 16, 128), torch.randn(2, 6, 128)
    mask = torch.tensor([[1, 1, 1, 0, 0, 0], [1, 0, 0, 0, 0, 0]])
    bias = ((1 - mask.float()) * -10000.0).unsqueeze(1)
    got = att(x, ctx, bias)
    q = att.to_q(x).view(2, 16, 4, 32).transpose(1, 2)
    k = att.to_k(ctx).view(2, 6, 4, 32).transpose(1, 2)
    v = att.to_v(ctx).view(2, 6, 4, 32).transpose(1, 2)
    ref = F.scaled_dot_product_attention(q, k, v, attn_mask=bias[:, None].expand(2, 4, 16, 6))
    ref = att.to_out[0](ref.transpose(1, 2).reshape(2, 16, 128))
    assert torch.allclose(got, ref, atol=1e-5)


def test_block_is_the_adaln_single_recipe():
    torch.manual_seed(1)
    cfg = DiTConfig(num_attention_heads=2, attention_head_dim=16, num_layers=1, sample_size=8)
    blk = dit.BasicTransformerBlock(cfg)
    x, ctx, tm = torch.randn(2, 16, 32), torch.randn(2, 5, 32), torch.randn(2, 6 * 32)
    got = blk(x, ctx, None, tm)
    sh1, sc1, g1, sh2, sc2, g2 = (blk.scale_shift_table[None] + tm.reshape(2, 6, 32)).unbind(1)
    ln = lambda z: F.layer_norm(z, (32,), eps=1e-6)
    y = x + g1[:, None] * blk.attn1(ln(x) * (1 + sc1[:, None]) + sh1[:, None])
    y = y + blk.attn2(y, ctx, None)
    y = y + g2[:, None] * blk.ff(ln(y) * (1 + sc2[:, None]) + sh2[:, None])
    assert torch.allclose(got, y, atol=1e-5)


def test_unpatchify_order_and_output_shape():
    torch.manual_seed(2)
    m = OracleDiT(TINY)
    lat = torch.randn(1, 4, 32, 32)
    out = m(lat, timestep=torch.tensor([10]), encoder_hidden_states=torch.randn(1, 3, 4096),
            encoder_attention_mask=torch.ones(1, 3, dtype=torch.long))[0]
    assert out.shape == (1, 8, 32, 32)
    # column (p*2+q)*8+c of proj_out lands at pixel (2h+p, 2w+q) of channel c
    y = torch.arange(16 * 16 * 32, dtype=torch.float32).reshape(1, 16, 16, 2, 2, 8)
    img = torch.einsum("nhwpqc->nchpwq", y).reshape(1, 8, 32, 32)
    assert img[0, 3, 2 * 5 + 1, 2 * 7 + 0] == y[0, 5, 7, 1, 0, 3]


def test_lora_is_peft_formula_and_starts_as_identity():
    from oracle import lora
    torch.manual_seed(3)
    m, m2 = OracleDiT(TINY), OracleDiT(TINY)
    m2.load_state_dict(m.state_dict())
    lora.add_lora(m2, r=64, alpha=64)
    names = [n for n, _ in m2.named_parameters() if ".lora_" in n]
    assert len(names) == 2 * 10 * TINY.num_layers and all(n.endswith(".default.weight") for n in names)
    assert all(p.requires_grad == (".lora_" in n) for n, p in m2.named_parameters())
    args = dict(timestep=torch.tensor([3, 700]), encoder_hidden_states=torch.randn(2, 4, 4096),
                encoder_attention_mask=torch.ones(2, 4, dtype=torch.long))
    lat = torch.randn(2, 4, 32, 32)
    with torch.no_grad():
        assert torch.equal(m(lat, **args)[0], m2(lat, **args)[0])  # B = 0
