"""TAESD decoder (reference main.py:166, 338, training.py:113-117).

CPU: the oracle restatement of diffusers' DecoderTiny (structure, parameter names, a comparison with the real class
wherever `diffusers` imports).  GPU: tdm_b200.taesd.AutoencoderTiny (tcgen05 implicit-GEMM convolutions, bf16 activations)
against the fp32 oracle on the same weights.  Tolerance: max|a-b| / max|b| <= 2e-2 and mean|a-b| / mean|b| <= 1e-2 (19
convolutions deep in bf16 against fp32)."""
import pytest
import torch

from oracle.taesd import OracleAutoencoderTiny, validation_images


def test_oracle_structure_and_names():
    torch.manual_seed(0)
    vae = OracleAutoencoderTiny()
    names = dict(vae.named_parameters())
    assert names["decoder.layers.0.weight"].shape == (64, 4, 3, 3)
    assert names["decoder.layers.18.weight"].shape == (3, 64, 3, 3) and names["decoder.layers.18.bias"].shape == (3,)
    for li in (6, 11, 16):  # convolution after each upsample: no bias
        assert names[f"decoder.layers.{li}.weight"].shape == (64, 64, 3, 3) and f"decoder.layers.{li}.bias" not in names
    blocks = [2, 3, 4, 7, 8, 9, 12, 13, 14, 17]
    for li in blocks:
        for ci in (0, 2, 4):
            assert names[f"decoder.layers.{li}.conv.{ci}.weight"].shape == (64, 64, 3, 3)
    assert len(names) == 2 + 10 * 6 + 3 + 2
    x = torch.randn(2, 4, 8, 8)
    out = vae.decode(x, return_dict=False)[0]
    assert out.shape == (2, 3, 64, 64)
    img = validation_images(vae, x)
    assert img.min() >= 0 and img.max() <= 1
    # x -> tanh(x / 3) * 3 clamps the latents: huge latents decode like +-3
    big = vae.decode(torch.full((1, 4, 8, 8), 1e4))[0]
    three = vae.decode(torch.full((1, 4, 8, 8), 3.0 * 1.0))[0]
    assert torch.isfinite(big).all() and (big - vae.decode(torch.full((1, 4, 8, 8), 50.0))[0]).abs().max() < 1e-4
    assert (big - three).abs().max() > 1e-4  # tanh(1) != 1


def test_oracle_against_real_diffusers_when_importable():
    diffusers = pytest.importorskip("diffusers")
    torch.manual_seed(0)
    ref = diffusers.AutoencoderTiny()
    ours = OracleAutoencoderTiny()
    missing, unexpected = ours.load_state_dict({k: v for k, v in ref.state_dict().items() if k.startswith("decoder.")}, strict=False)
    assert not missing
    x = torch.randn(2, 4, 8, 8)
    with torch.no_grad():
        assert torch.allclose(ref.decode(x, return_dict=False)[0], ours.decode(x)[0], atol=1e-5)


def _pair(seed=0):
    from tdm_b200.taesd import AutoencoderTiny
    torch.manual_seed(seed)
    ref = OracleAutoencoderTiny()
    with torch.no_grad():  # biases of a trained TAESD are not tiny: make them matter
        for n, p in ref.named_parameters():
            if n.endswith(".bias"):
                p.mul_(3.0)
        ref.decoder.layers[18].bias.fill_(0.5)   # centre the output: x * 2 - 1 around 0, so images are not saturated
        ref.decoder.layers[18].weight.mul_(4.0)
    ours = AutoencoderTiny(device="cuda", init_weights=False)
    ours.load_state_dict(ref.state_dict())
    return ref, ours


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / b.abs().max()).item(), ((a - b).abs().mean() / b.abs().mean()).item()


@pytest.mark.gpu
@pytest.mark.timeout(600)
@pytest.mark.parametrize("B,H,W", [(2, 8, 8), (1, 16, 16), (3, 32, 32), (1, 64, 64), (2, 16, 32)])
def test_decode_matches_oracle(B, H, W):
    ref, ours = _pair()
    x = torch.randn(B, 4, H, W, generator=torch.Generator().manual_seed(B * 100 + H)) * 2.0
    with torch.no_grad():
        want = ref.decode(x)[0]
    got = ours.decode(x.cuda(), return_dict=False)[0]
    assert got.shape == want.shape == (B, 3, 8 * H, 8 * W) and got.dtype == torch.float32
    mx, mean = _rel(got, want)
    assert mx < 2e-2 and mean < 1e-2, (mx, mean)


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_decode_variants_and_validation_images():
    from tdm_b200 import harness
    from tdm_b200.taesd import decode_to_images
    from tdm_b200.training import save_validation_images
    ref, ours = _pair(seed=1)
    x = torch.randn(5, 4, 16, 16, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        want = validation_images(ref, x)
    ours.max_images_per_launch = 2           # 5 images in launches of 2 + 2 + 1
    got = decode_to_images(ours, x.cuda())
    assert want.std() > 0.05                  # not a saturated image: the comparison means something
    assert (got.cpu() - want).abs().max().item() < 2e-2 and got.min() >= 0 and got.max() <= 1
    ours.max_images_per_launch = 16
    again = decode_to_images(ours, x.cuda())
    assert torch.equal(got, again)           # chunking does not change a bit
    for dt in (torch.bfloat16, torch.float16):
        with torch.no_grad():
            want_dt = ref.decode(x.to(dt).float())[0]
        assert _rel(ours.decode(x.to(dt).cuda())[0], want_dt)[0] < 2e-2
    with pytest.raises(Exception, match="CUDA"):
        ours.decode(x)
    # training.py:107-121 on a tiny student: 4-step and 1-step samples decoded to [0, 1] images
    student, _, _ = harness.make_models("tiny", seed=0, device="cuda", lora_rank=0, teacher=False, fake=False)
    ctx, mask, _, _ = harness.synthetic_text(4, L=24, seed=0)
    noise = torch.randn(4, 4, 32, 32).to(torch.bfloat16).cuda()
    img4, img1 = save_validation_images(student, harness.make_scheduler(), ours, ctx.cuda(), mask.cuda(), noise, 900,
                                        lowp=torch.bfloat16)
    assert img4.shape == img1.shape == (4, 3, 256, 256)
    assert torch.isfinite(img4).all() and img4.min() >= 0 and img4.max() <= 1
