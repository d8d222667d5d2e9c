"""Data-parallel equivalence on real GPUs over NCCL: 2 ranks x batch B must produce the
gradients and the post-step weights of 1 rank x batch 2B.

The path shards by batch: no op couples samples, the TDM weight w_b is per sample (main.py:519), both losses are
means, so summed per-rank gradients x 1/world equal the global-batch gradients up to fp32 reduction order
(different split of the token reductions, atomics) -- the bf16 operands of every GEMM row are identical.
Skipped with fewer than 2 GPUs."""
import os
import sys
from pathlib import Path

import pytest
import torch
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]

B_RANK, WORLD, LR = 2, 2, 1e-3


def _inputs(Cc, S, L=24):
    """Global batch (WORLD * B_RANK samples), drawn on the CPU generator."""
    from tdm_b200 import harness
    from tdm_b200.training import draw_fake_branch
    B = WORLD * B_RANK
    ctx, mask, uctx, umask = harness.synthetic_text(B, L=L, Cc=Cc, seed=3)
    torch.manual_seed(17)
    noise = torch.randn(B, 4, S, S).to(torch.bfloat16)
    ind_t, t_fake, noise_fake = draw_fake_branch(noise, 900, True)
    ind_t = torch.tensor([1, 2, 3, 4])[:B]  # every trajectory state is selected by some sample
    return noise, ctx, mask, uctx, umask, (ind_t, t_fake, noise_fake)


def _run(device, rows, grad_sync):
    """One TdmTrainer.step on `rows` of the global batch; returns CPU copies of (losses, grads before the optimiser
    step, weight deltas of the step)."""
    from tdm_b200 import harness
    from tdm_b200.training import TdmTrainer
    student, fake, teacher = harness.make_models("tiny", seed=0, device=device, lora_rank=64)
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():  # LoRA B = 0 would zero dA; give the fake model a different head so the loss is conditioned
        for n, p in student.named_parameters():
            if ".lora_B." in n:
                p.copy_((torch.randn(p.shape, generator=g) * 0.02).to(device))
        dict(fake.named_parameters())["proj_out.weight"].mul_(1.3)
    noise, ctx, mask, uctx, umask, (ind_t, t_fake, noise_fake) = _inputs(student.Cc, student.sample_size)
    sl = slice(*rows)
    tr = TdmTrainer(student, fake, teacher, harness.make_scheduler(), uctx.to(device), umask.to(device), lr=LR, lr_fake=LR,
                    lowp=torch.bfloat16, grad_sync=grad_sync)
    kept = {}
    for name, opt in (("student", tr.opt_student), ("fake", tr.opt_fake)):
        zero = opt.zero_grad

        def capture(name=name, opt=opt, zero=zero):
            kept[name] = torch.cat([g_.detach().flatten().clone() for _, g_, _ in opt.groups])
            zero()
        opt.zero_grad = capture
    w0 = {"student": student._lflat.clone(), "fake": fake._flat.clone()}
    out = tr.step(noise[sl].to(device), ctx[sl].to(device), mask[sl].to(device),
                  draws=(ind_t[sl].to(device), t_fake[sl].to(device), noise_fake[sl].to(device)))
    tr.sync()
    torch.cuda.synchronize(device)
    world = grad_sync.world_size if grad_sync is not None else 1
    return {"loss": out["loss"].item(), "loss_fake": out["loss_fake"].item(),
            "g_student": (kept["student"] / world).cpu(), "g_fake": (kept["fake"] / world).cpu(),
            "d_student": (student._lflat - w0["student"]).cpu(), "d_fake": (fake._flat - w0["fake"]).cpu(),
            "norm_student": tr.opt_student.grad_norm.item(), "norm_fake": tr.opt_fake.grad_norm.item()}


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    import torch.distributed as dist
    from tdm_b200 import dp
    dp.init_from_env("nccl")
    dev = torch.device("cuda", rank)
    res = _run(dev, (rank * B_RANK, (rank + 1) * B_RANK), dp.GradSync())
    q.put((rank, res))
    dist.barrier(device_ids=[rank])
    dist.destroy_process_group()


def _rel(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def test_two_ranks_equal_one_rank_with_the_global_batch():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (NCCL)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, WORLD, port, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ref = _run(torch.device("cuda", 0), (0, WORLD * B_RANK), None)
    r0, r1 = got[0], got[1]
    # after the all-reduce every rank holds the same summed gradients and takes the same step
    for k in ("g_student", "g_fake", "d_student", "d_fake"):
        assert torch.equal(r0[k], r1[k]), k
    # losses: mean over ranks of the per-rank means == global-batch mean
    for k in ("loss", "loss_fake"):
        avg = 0.5 * (r0[k] + r1[k])
        assert abs(avg - ref[k]) <= 1e-5 * abs(ref[k]), (k, avg, ref[k])
    # gradients: fp32 reduction noise only (global max norm AND mean absolute error against the mean magnitude)
    for k in ("g_student", "g_fake"):
        assert ref[k].abs().max() > 0
        assert _rel(r0[k], ref[k]) < 2e-3, (k, _rel(r0[k], ref[k]))
        assert ((r0[k] - ref[k]).abs().mean() / ref[k].abs().mean()).item() < 2e-3, k
    for k in ("norm_student", "norm_fake"):
        assert abs(r0[k] - ref[k]) <= 1e-3 * ref[k], (k, r0[k], ref[k])
    # post-step weights: AdamW's first step moves every weight by ~lr * sign(g); only entries whose gradient is
    # rounding noise may differ, and never by more than 2 lr
    for k in ("d_student", "d_fake"):
        diff = (r0[k] - ref[k]).abs()
        assert diff.max().item() <= 2.001 * LR, k
        moved = ref[k].abs() > 0
        assert moved.float().mean().item() > 0.5, k
        assert (diff > 0.1 * LR).float().mean().item() < 1e-2, (k, (diff > 0.1 * LR).float().mean().item())
        assert (diff.mean() / ref[k].abs().mean()).item() < 1e-2, k
