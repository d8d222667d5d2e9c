#!/usr/bin/env python
"""bench.py -- TDM distillation hot path on B200 (BASELINE.json metric: distill iters/s + 4-NFE img/s,
PixArt-alpha-512, 1/2/4/8 GPUs vs the bf16 tensor-core roofline).

    python bench.py --gpus 1 --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus 1 ...            # the reference's algorithm on the host CPU cores
    torchrun --nproc-per-node N bench.py --gpus N ...        # one rank per GPU, NCCL gradient all-reduce

One step = one TDM distillation iteration (DESIGN.md "The iteration") on one batch of synthetic input
(batch 8 per GPU, PixArt-alpha-512 DiT-XL/2, 120-token synthetic T5 context, random-init weights, bf16).
Rank 0 prints ONE JSON line.
  value   : whole-job throughput, inputs resident in HBM: batch-8 iterations per second summed over ranks.
  e2e     : the same through TdmTrainer.step_from_host -- pinned HOST buffers in, host loss out, every step.
  roofline: dominant kernel (tcgen05 GEMM) achieved TFLOP/s timed live with CUDA events, against
            MEASURED_PEAKS.json; `step_roofline` is the whole iteration against the sustained peak.
  cpu_baseline: the oracle (a port: plain PyTorch fp32 restatement) timed on the host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

METRIC = "tdm_distill_iters_per_s"
UNIT = "batch-8 iterations/s, summed over GPUs (PixArt-alpha-512 DiT-XL/2, 4-NFE student trajectory)"


# ---------------------------------------------------------------------------------------------------
# FLOP accounting (2MNK per GEMM, 4*Sq*Sk*D per attention, no padding, no recompute)
# ---------------------------------------------------------------------------------------------------
def dit_flops(cfg, L=120, lora_r=0):
    """Algorithmic FLOPs of ONE sample: forward, and the extra terms of the two backward flavours."""
    D = cfg["num_attention_heads"] * cfg["attention_head_dim"]
    N = (cfg["sample_size"] // cfg["patch_size"]) ** 2
    depth, Cc, p = cfg["num_layers"], cfg["caption_channels"], cfg["patch_size"]
    lin_block = 2 * N * D * D * (3 + 1 + 1 + 1) + 2 * L * D * D * 2 + 2 * N * D * 4 * D * 2  # = 28 N D^2 + 4 L D^2
    attn_block = 4 * N * N * D + 4 * N * L * D
    other = (2 * N * (cfg["in_channels"] * p * p) * D + 2 * (256 * D + D * D + D * 6 * D)
             + 2 * L * (Cc * D + D * D) + 2 * N * D * (p * p * cfg["out_channels"]))
    fwd = depth * (lin_block + attn_block) + other
    # LoRA side products: sum over the 10 adapted linears of tokens*(in+out):
    # q,k,v,o,q2,o2: N*2D each; k2,v2: L*2D each; ff1, ff2: N*5D each
    lora_terms = 22 * N * D + 4 * L * D
    lora_fwd = depth * 2 * lora_r * lora_terms if lora_r else 0  # t = x A^T and t B^T
    attn_bwd = 2.5 * attn_block  # 5 contractions (S, dP, dV, dK, dQ) vs 2 forward
    # full fine-tune backward: dgrad + wgrad of every linear + attention backward
    bwd_full = depth * (2 * lin_block + attn_bwd) + 2 * other
    # LoRA student backward: dgrad of every linear (input gradients must flow), no base wgrad, + LoRA terms
    # (dt = dy B, dB = dy^T t, dA = dt^T x, and dt A folded into the dgrad)
    bwd_lora = depth * (lin_block - 2 * L * D * D * 2 + attn_bwd) + (depth * 4 * lora_r * lora_terms if lora_r else 0) \
        + 2 * N * D * (p * p * cfg["out_channels"])
    return dict(fwd=fwd, lora_fwd=lora_fwd, bwd_full=bwd_full, bwd_lora=bwd_lora, lin_block=lin_block, attn_block=attn_block)


def iteration_flops(cfg, B, lora_r):
    f = dit_flops(cfg, lora_r=lora_r)
    passes = {"student_fwd_nograd": 4, "student_fwd_grad": 1, "student_bwd_lora": 1, "fake_fwd": 1, "fake_bwd_full": 1,
              "teacher_fwd": 2}
    per_sample = (5 * (f["fwd"] + f["lora_fwd"]) + f["bwd_lora"] + f["fwd"] + f["bwd_full"] + 2 * f["fwd"])
    return per_sample * B, passes, f


# ---------------------------------------------------------------------------------------------------
# clocks sampling during the timed region (nvidia-smi fields of B200_PROFILING.md, through NVML)
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index=0, period=0.2):
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.sm_max = None
        self._stop = threading.Event()
        self._th = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv = None
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        if self.nv is None:
            return self
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()
        return self

    def stop(self):
        if self._th is not None:
            self._stop.set()
            self._th.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(s)}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d["hbm_gbs"], burst=d["bf16_tflops"], sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="MEASURED_PEAKS.json (of measured)")
    return dict(hbm=6650.0, burst=1590.0, sustained=1400.0, source="B200_PROFILING.md fallback (of fallback)")


# ---------------------------------------------------------------------------------------------------
# reference arm: the reference's algorithm (oracle port, plain PyTorch fp32) on the host cores
# ---------------------------------------------------------------------------------------------------
def build_oracle(config_name, lora_r, seed=0):
    from oracle.dit import OracleDiT, DiTConfig
    from oracle import lora as oracle_lora
    from oracle.schedule import OracleDDPMScheduler
    from tdm_b200.harness import CONFIGS
    cfg = DiTConfig(**CONFIGS[config_name])
    torch.manual_seed(seed)
    student = OracleDiT(cfg)
    fake, teacher = OracleDiT(cfg), OracleDiT(cfg)
    sd = student.state_dict()
    fake.load_state_dict(sd)
    teacher.load_state_dict(sd)
    if lora_r:
        oracle_lora.add_lora(student, r=lora_r, alpha=lora_r)
    return cfg, student, fake, teacher, OracleDDPMScheduler()


def cpu_iteration_seconds(config_name, B_sample, lora_r, steps, warmup, threads=None):
    """Times oracle.iteration.RepairedIteration.step at batch B_sample on the host cores."""
    from oracle.iteration import RepairedIteration
    from tdm_b200.harness import synthetic_text
    if threads:
        torch.set_num_threads(threads)
    cfg, student, fake, teacher, sch = build_oracle(config_name, lora_r)
    ctx, mask, uctx, umask = synthetic_text(B_sample, seed=0, Cc=cfg.caption_channels)
    it = RepairedIteration(student, fake, teacher, sch, uctx, umask, lowp=torch.bfloat16)
    S = cfg.sample_size
    times = []
    for i in range(warmup + steps):
        torch.manual_seed(100 + i)
        noise = torch.randn(B_sample, 4, S, S).to(torch.bfloat16)
        t0 = time.perf_counter()
        it.step(noise, ctx, mask)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times), torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    B = args.batch
    B_sample = 1
    sec, cores = cpu_iteration_seconds(args.config, B_sample, args.lora_rank, max(1, min(args.steps, 2)), 0)
    iters_per_s = 1.0 / (sec * (B / B_sample))
    sample = (f"oracle/iteration.RepairedIteration.step (plain PyTorch fp32 port of the reference's algorithm) at batch "
              f"{B_sample} of {B}, full {args.config} model, {max(1, min(args.steps, 2))} timed iteration(s); batch-{B} rate = 1/({B}/{B_sample} x t)")
    line = {"impl": "reference", "metric": METRIC, "value": iters_per_s, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * (B / B_sample) * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"TDM distillation iteration, {args.config}, batch {B}, 120-token synthetic T5 ctx, LoRA r={args.lora_rank} student"},
            "cpu_baseline": {"value": iters_per_s, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": iters_per_s, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def gemm_roofline(cfg, B, peaks, iters=20):
    """Live timing of the dominant kernel (gemm2_bf16_kernel, the CTA-pair GEMM) on the block's forward GEMM shapes, CUDA events on
    the launching stream; achieved = algorithmic FLOPs / time."""
    from tdm_b200 import ops
    D = cfg["num_attention_heads"] * cfg["attention_head_dim"]
    N = (cfg["sample_size"] // cfg["patch_size"]) ** 2
    M = B * N
    shapes = [(M, 3 * D, D), (M, D, D), (M, D, D), (M, D, D), (M, 4 * D, D), (M, D, 4 * D)]
    dev = "cuda"
    bufs = []
    for (m, n, k) in shapes:
        bufs.append((torch.randn(m, k, device=dev).bfloat16(), torch.randn(n, k, device=dev).bfloat16(),
                     torch.empty(m, n, device=dev, dtype=torch.bfloat16)))
    flops = sum(2.0 * m * n * k for m, n, k in shapes)
    for a, b, c in bufs:
        ops.gemm(a, b, out_bf16=c)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        for a, b, c in bufs:
            ops.gemm(a, b, out_bf16=c)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    ach = flops / (ms * 1e-3) / 1e12
    # DRAM traffic per launch from the committed `ncu --set full` capture (profiles/r01_ncu_full_v9_summary.txt):
    # FF1 shape 8192x4608x1152 (bias + GELU-tanh, bf16 out), dram__bytes_read.sum + dram__bytes_write.sum = 29.59 + 31.31 MB
    # (algorithmic operand + output bytes 105 MB: part of the bf16 output is still in the 126 MB L2 when the kernel ends,
    # nothing is re-read).
    return {"bound": "tensor", "kernel": "gemm2_bf16_kernel<256|192, fast> (CTA-pair tcgen05/TMEM/TMA GEMM, streaming TMA epilogue; half of the iteration), the 6 forward GEMM shapes of one DiT block at M=%d" % M,
            "achieved": ach, "peak": peaks["burst"], "unit": "TFLOP/s", "frac": ach / peaks["burst"],
            "traffic": {"bytes_per_launch": 60.9e6, "shape": "8192x4608x1152 (FF1, bias + GELU-tanh)", "source": "ncu --set full, profiles/r01_ncu_full_v9_summary.txt"},
            "peak_source": peaks["source"] + ", burst figure (kernel timed alone)", "launches_timed": iters * len(shapes),
            "avg_launch_us": ms * 1e3 / len(shapes)}


def run_ours(args):
    from tdm_b200 import _lib, dp, harness
    from tdm_b200.training import TdmTrainer, draw_iteration_noise, sample_4nfe, sample_4nfe_from_host
    import torch.distributed as dist
    rank, local, world = dp.init_from_env("nccl")
    if world != args.gpus and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    lib = _lib.lib()
    peaks = measured_peaks()
    cfg = harness.CONFIGS[args.config]
    B, lowp = args.batch, torch.bfloat16
    S = cfg["sample_size"]
    student, fake, teacher = harness.make_models(args.config, seed=0, device=dev, lora_rank=args.lora_rank)
    dp.broadcast_params([student, fake, teacher])
    gs = dp.GradSync() if world > 1 else None
    ctx_h, mask_h, uctx_h, umask_h = harness.synthetic_text(B, seed=1000 + rank, Cc=cfg["caption_channels"], pin=True)
    ctx_h = ctx_h.to(torch.bfloat16).pin_memory()  # T5 embeddings are handed over in the compute dtype
    trainer = TdmTrainer(student, fake, teacher, harness.make_scheduler(), uctx_h.to(dev).to(torch.bfloat16), umask_h.to(dev),
                         cfg=4.5, total_steps=900, use_huber=True, use_separate=True, lowp=lowp, grad_sync=gs,
                         defer_fake_opt=os.environ.get("TDM_B200_DEFER_OPT", "1") != "0")
    torch.manual_seed(1234 + rank)  # per-rank data stream (accelerate's set_seed seeds ranks identically)
    n_total = args.warmup + 2 * args.steps
    noises_h = [draw_iteration_noise(B, (S, S), lowp, "cpu").pin_memory() for _ in range(n_total)]
    ctx_d, mask_d = ctx_h.to(dev), mask_h.to(dev)
    noises_d = [n.to(dev) for n in noises_h]

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    # ---- warm-up ----
    for i in range(args.warmup):
        trainer.step(noises_d[i], ctx_d, mask_d)
    barrier()
    # ---- timed region A: inputs resident in HBM ----
    sampler = ClockSampler(local).start() if rank == 0 else None
    lib.tdm_launch_count(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        out = trainer.step(noises_d[args.warmup + i], ctx_d, mask_d)
    trainer.sync()  # a deferred fake-score optimiser step belongs to the timed region
    e1.record()
    barrier()
    launches = int(lib.tdm_launch_count(0))
    ms_a = e0.elapsed_time(e1)
    # ---- timed region B: end to end from pinned host buffers, host loss out ----
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e2.record()
    for i in range(args.steps):
        loss_h, loss_fake_h = trainer.step_from_host(noises_h[args.warmup + args.steps + i], ctx_h, mask_h)
    trainer.sync()
    e3.record()
    barrier()
    ms_b = e2.elapsed_time(e3)
    clocks = sampler.stop() if sampler is not None else None
    t = torch.tensor([ms_a, ms_b], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_a, ms_b = t.tolist()
    h2d = noises_h[0].numel() * noises_h[0].element_size() + ctx_h.numel() * ctx_h.element_size() + mask_h.numel() * mask_h.element_size()
    d2h = 16

    # ---- sampling: 4-NFE images/s (BASELINE.json configs[3]) ----
    sampling = None
    if args.sample_batch > 0:
        Bs = args.sample_batch
        sctx_h, smask_h, _, _ = harness.synthetic_text(Bs, seed=2000 + rank, Cc=cfg["caption_channels"], pin=True)
        sctx_h = sctx_h.to(torch.bfloat16).pin_memory()
        snoise_h = draw_iteration_noise(Bs, (S, S), lowp, "cpu").pin_memory()
        sctx_d, smask_d, snoise_d = sctx_h.to(dev), smask_h.to(dev), snoise_h.to(dev)
        sched = harness.make_scheduler()
        for _ in range(2):
            sample_4nfe(student, sched, snoise_d, sctx_d, smask_d, lowp)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(2, args.steps)
        s0.record()
        for _ in range(reps):
            sample_4nfe(student, sched, snoise_d, sctx_d, smask_d, lowp)
        s1.record()
        barrier()
        ms_s = s0.elapsed_time(s1) / reps
        s2, s3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s2.record()
        for _ in range(reps):
            sample_4nfe_from_host(student, sched, snoise_h, sctx_h, smask_h, lowp)
        s3.record()
        barrier()
        ms_se = s2.elapsed_time(s3) / reps
        ts = torch.tensor([ms_s, ms_se], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        ms_s, ms_se = ts.tolist()
        f = dit_flops(cfg, lora_r=args.lora_rank)
        fl = 4 * (f["fwd"] + f["lora_fwd"]) * Bs
        sampling = {"metric": "images_per_s_4nfe", "value": world * Bs / (ms_s * 1e-3), "e2e": world * Bs / (ms_se * 1e-3),
                    "unit": "final latents/s (4 NFE, timesteps 999/856/665/399), summed over GPUs", "batch_per_gpu": Bs,
                    "ms_per_batch": ms_s, "tflops_per_gpu": fl / (ms_s * 1e-3) / 1e12,
                    "frac_of_sustained_peak": fl / (ms_s * 1e-3) / 1e12 / peaks["sustained"]}

    if rank != 0:
        if world > 1:
            dist.barrier(device_ids=[local])
            dist.destroy_process_group()
        return 0

    ms_step = ms_a / args.steps
    value = world / (ms_step * 1e-3)
    e2e = world / (ms_b / args.steps * 1e-3)
    flops_iter, passes, f = iteration_flops(cfg, B, args.lora_rank)
    step_tf = flops_iter / (ms_step * 1e-3) / 1e12
    roof = gemm_roofline(cfg, B, peaks)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": f"TDM distillation iteration (BASELINE.json configs[{1 if world == 1 else 2}]): {args.config} DiT-XL/2 "
                               f"(28 blocks, 1152 hidden, 64x64x4 latent, 120-token synthetic T5 ctx), batch {B}/GPU, random-init, "
                               f"LoRA r={args.lora_rank} student + full fake-score model + frozen teacher, cfg 4.5, Huber, --use_separate",
                   "passes_per_sample": passes, "batch_per_gpu": B, "parallelism": f"dp{world}",
                   "l2": "no explicit flush: each step streams 3 x 1.2 GB of bf16 weights + GBs of activations, far above the 126 MB L2"},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_b / args.steps},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roof,
        "step_roofline": {"bound": "tensor", "achieved": step_tf, "peak": peaks["sustained"], "unit": "TFLOP/s",
                          "frac": step_tf / peaks["sustained"], "flops_per_step_per_gpu": flops_iter,
                          "peak_source": peaks["source"] + ", sustained figure (timed inside a long step)"},
        "sampling": sampling,
        "loss": {"tdm": loss_h, "fake": loss_fake_h},
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            sec, cores = cpu_iteration_seconds("pixart-512" if args.config == "pixart-512" else args.config, 1, args.lora_rank, 1, 0)
            v = 1.0 / (sec * B)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"one oracle iteration (plain PyTorch fp32) at batch 1 of {B}, full model: {sec:.1f} s; batch-{B} rate = 1/({B} x t)"}
        except Exception as e:  # the baseline is reported, never required for the GPU numbers
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": f"failed: {e!r}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(device_ids=[local])
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--config", default="pixart-512", choices=["tiny", "pixart-512", "pixart-1024"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--lora-rank", type=int, default=64)
    ap.add_argument("--sample-batch", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
