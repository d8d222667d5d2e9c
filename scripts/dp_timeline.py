"""Timeline of the data-parallel gradient exchange (nsys is not in the image: torch.profiler / CUPTI instead).
    torchrun --nproc-per-node N scripts/dp_timeline.py [out.json]
Rank 0 profiles ONE iteration after warm-up and writes, per NCCL all-reduce kernel: start, duration, how much of it ran
under this library's compute kernels (any stream), and the exposed tail after the last compute kernel."""
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    import torch.distributed as dist
    from torch.profiler import ProfilerActivity, profile
    from tdm_b200 import dp, harness
    from tdm_b200.training import TdmTrainer, draw_iteration_noise
    out_path = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "gpurun_out" / "dp_timeline.json")
    config = os.environ.get("TDM_B200_CONFIG", "pixart-512")
    rank, local, world = dp.init_from_env("nccl")
    dev = torch.device("cuda", local)
    cfg = harness.CONFIGS[config]
    S, B = cfg["sample_size"], 8
    student, fake, teacher = harness.make_models(config, seed=0, device=dev, lora_rank=64)
    dp.broadcast_params([student, fake, teacher])
    gs = dp.GradSync() if world > 1 else None
    ctx, mask, uctx, umask = harness.synthetic_text(B, seed=1000 + rank, Cc=cfg["caption_channels"])
    ctx, mask = ctx.to(dev).to(torch.bfloat16), mask.to(dev)
    tr = TdmTrainer(student, fake, teacher, harness.make_scheduler(), uctx.to(dev).to(torch.bfloat16), umask.to(dev), grad_sync=gs)
    torch.manual_seed(1234 + rank)
    noises = [draw_iteration_noise(B, (S, S), torch.bfloat16, dev) for _ in range(5)]
    for i in range(3):
        tr.step(noises[i], ctx, mask)
    tr.sync()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(device_ids=[local])
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        tr.step(noises[3], ctx, mask)
        tr.sync()
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier(device_ids=[local])
    if rank == 0:
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
        ker = [(e.name, e.time_range.start, e.time_range.end) for e in evs if "memcpy" not in e.name.lower() and "memset" not in e.name.lower()]
        t0 = min(s for _, s, _ in ker)
        nccl = sorted([(n, s - t0, e - t0) for n, s, e in ker if "nccl" in n.lower()], key=lambda x: x[1])
        comp = sorted([(n, s - t0, e - t0) for n, s, e in ker if "nccl" not in n.lower()], key=lambda x: x[1])
        # merged busy intervals of the compute kernels
        merged = []
        for _, s, e in comp:
            if merged and s <= merged[-1][1]:
                merged[-1][1] = max(merged[-1][1], e)
            else:
                merged.append([s, e])
        def overlap(s, e):
            return sum(max(0.0, min(e, b) - max(s, a)) for a, b in merged)
        last_comp = max(e for _, _, e in comp)
        last_bwd = max((e for n, _, e in comp if "adamw" not in n and "sumsq" not in n and "clip" not in n and "memset" not in n.lower()), default=last_comp)
        first_opt = min((s for n, s, _ in comp if "sumsq" in n or "adamw" in n), default=None)
        recs = [{"i": i, "start_us": round(s, 1), "dur_us": round(e - s, 1), "under_compute_us": round(overlap(s, e), 1)} for i, (n, s, e) in enumerate(nccl)]
        summ = {"world": world, "config": config, "iteration_us": round(max(e for _, _, e in ker), 1), "n_compute_kernels": len(comp),
                "n_nccl_kernels": len(nccl), "nccl_total_us": round(sum(r["dur_us"] for r in recs), 1),
                "nccl_under_compute_us": round(sum(r["under_compute_us"] for r in recs), 1),
                "nccl_exposed_us": round(sum(r["dur_us"] - r["under_compute_us"] for r in recs), 1),
                "first_optimizer_kernel_start_us": None if first_opt is None else round(first_opt, 1),
                "last_nccl_end_us": round(max((e for _, _, e in nccl), default=0.0), 1),
                "gap_before_optimizer_us": None if first_opt is None else round(first_opt - max([e for n, s, e in comp if e <= first_opt] + [0.0]), 1),
                "allreduce": recs}
        Path(out_path).parent.mkdir(exist_ok=True)
        Path(out_path).write_text(json.dumps(summ, indent=1))
        print(json.dumps({k: v for k, v in summ.items() if k != "allreduce"}))
    if world > 1:
        dist.barrier(device_ids=[local])
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
