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
    # the 4 trajectory forwards run under no_grad on MERGED weights (W_eff = W + s B A): no LoRA side products there;
    # only the activations-kept student forward executes x A^T / t B^T
    per_sample = (4 * f["fwd"] + (f["fwd"] + f["lora_fwd"]) + f["bwd_lora"] + f["fwd"] + f["bwd_full"] + 2 * f["fwd"])
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


def host_threads():
    """Threads the CPU legs may use: the cores this process is allowed on (torchrun exports OMP_NUM_THREADS=1, which
    would otherwise pin the oracle to one core)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def workload_string(config_name, cfg, B, lora_r, world=1):
    S = cfg["sample_size"]
    idx = {"tiny": 0, "pixart-512": 1 if world == 1 else 2, "pixart-1024": 4}.get(config_name, 1)
    return (f"TDM distillation iteration (BASELINE.json configs[{idx}]): {config_name} DiT ({cfg['num_layers']} blocks, "
            f"{cfg['num_attention_heads'] * cfg['attention_head_dim']} hidden, {S}x{S}x4 latent = {(S // cfg['patch_size']) ** 2} tokens, "
            f"120-token synthetic T5 ctx), batch {B}/GPU, random-init, LoRA r={lora_r} student + full fake-score model + "
            f"frozen teacher, cfg 4.5, Huber, --use_separate")


def cpu_iteration_seconds(config_name, B_sample, lora_r, steps, warmup, threads=None, warmup_batch=None, sdpa=True):
    """Times oracle.iteration.RepairedIteration.step at batch B_sample on the host cores (fp32, all threads).
    Warm-up iterations run at `warmup_batch` (default B_sample).  Returns (seconds per timed iteration, threads, n timed)."""
    from oracle import dit as oracle_dit
    from oracle.iteration import RepairedIteration
    from tdm_b200.harness import synthetic_text
    torch.set_num_threads(threads or host_threads())
    oracle_dit.Attention.use_sdpa = bool(sdpa)  # same function without materialising S (the reference runs xformers)
    try:
        cfg, student, fake, teacher, sch = build_oracle(config_name, lora_r)
        ctx, mask, uctx, umask = synthetic_text(B_sample, seed=0, Cc=cfg.caption_channels)
        it = RepairedIteration(student, fake, teacher, sch, uctx, umask, lowp=torch.bfloat16)
        S = cfg.sample_size
        times = []
        for i in range(warmup + steps):
            b = (warmup_batch or B_sample) if i < warmup else B_sample
            torch.manual_seed(100 + i)
            noise = torch.randn(b, 4, S, S).to(torch.bfloat16)
            t0 = time.perf_counter()
            it.step(noise, ctx[:b], mask[:b])
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    finally:
        oracle_dit.Attention.use_sdpa = False
    return sum(times) / len(times), torch.get_num_threads(), len(times)


def reference_batch(B, config_name):
    """Largest batch <= B whose fp32 autograd footprint fits the host: ~11 GB of saved activations per sample for two
    DiT-XL/2-512 forwards with grad (student + fake) + 3 x 2.4 GB weights + grads + Adam state."""
    try:
        import psutil
        avail = psutil.virtual_memory().available / 2 ** 30
    except Exception:
        avail = 32.0
    per_sample = {"tiny": 0.05, "pixart-512": 11.0, "pixart-1024": 45.0}.get(config_name, 11.0)
    fixed = {"tiny": 0.5, "pixart-512": 22.0, "pixart-1024": 22.0}.get(config_name, 22.0)
    b = B
    while b > 1 and fixed + per_sample * b > 0.6 * avail:
        b //= 2
    return b, avail


def run_reference(args):
    """The reference's algorithm on the box's host cores (oracle port; the reference itself is pure Python on top of
    diffusers / accelerate, which are not installable here -- DESIGN.md section 6).  Rank 0 only."""
    from tdm_b200.harness import CONFIGS
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    B = args.batch
    B_sample, avail = reference_batch(B, args.config)
    timed = max(1, min(args.steps, 2 if B_sample == B else 3))
    warm = 1 if args.warmup > 0 else 0
    sec, cores, n = cpu_iteration_seconds(args.config, B_sample, args.lora_rank, timed, warm, warmup_batch=1)
    scale = B / B_sample
    iters_per_s = 1.0 / (sec * scale)
    sample = (f"oracle/iteration.RepairedIteration.step (plain PyTorch fp32 port of the reference's algorithm, SDPA attention) on "
              f"{cores} host threads: {n} timed iteration(s) at batch {B_sample} of {B} after {warm} warm-up iteration at batch 1, "
              f"full {args.config} model, {sec:.1f} s each" + ("" if scale == 1 else f"; batch-{B} rate = 1/({scale:g} x t), host RAM {avail:.0f} GB free"))
    line = {"impl": "reference", "metric": METRIC, "value": iters_per_s, "unit": UNIT, "n_gpus": args.gpus, "steps": n,
            "warmup": warm, "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": sec * scale * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "same_config": scale == 1,
            "config": {"workload": workload_string(args.config, CONFIGS[args.config], B, args.lora_rank, args.gpus),
                       "batch_timed": B_sample, "batch_per_gpu": B},
            "cpu_baseline": {"value": iters_per_s, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": iters_per_s, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def gemm_roofline(cfg, B, peaks, device_index=0, min_ms=300.0):
    """Live timing of the dominant kernel (gemm2_bf16_kernel, the CTA-pair tcgen05 GEMM) on the six GEMMs of one DiT
    block's no-grad forward WITH the epilogues the iteration runs them with (bias->bf16; bias+gate+residual->fp32 and a
    bf16 copy; bias+residual->fp32; bias+GELU-tanh->bf16), CUDA events on the launching stream, clocks sampled around
    this region; achieved = algorithmic 2MNK / time."""
    from tdm_b200 import ops
    D = cfg["num_attention_heads"] * cfg["attention_head_dim"]
    N = (cfg["sample_size"] // cfg["patch_size"]) ** 2
    M = B * N
    dev = "cuda"
    bf = dict(device=dev, dtype=torch.bfloat16)
    f32 = dict(device=dev, dtype=torch.float32)
    act = lambda k: torch.randn(M, k, **f32).to(torch.bfloat16)
    wgt = lambda n, k: (torch.randn(n, k, **f32) * k ** -0.5).to(torch.bfloat16)
    h1, a1, x1b, a2, h2 = act(D), act(D), act(D), act(D), act(D)
    g = act(4 * D)
    x, x1, x2 = torch.randn(M, D, **f32), torch.randn(M, D, **f32), torch.randn(M, D, **f32)
    mod = torch.randn(B, 6 * D, **f32)
    bias = lambda n: torch.randn(n, **f32)
    calls = [
        ("qkv  bias->bf16", (M, 3 * D, D), dict(a=h1, b=wgt(3 * D, D), bias=bias(3 * D), out_bf16=torch.empty(M, 3 * D, **bf))),
        ("o    bias+gate+res->f32+bf16", (M, D, D), dict(a=a1, b=wgt(D, D), bias=bias(D), gate=mod[:, 2 * D:3 * D], rows_per_gate=N, res=x,
                                                        out_f32=torch.empty(M, D, **f32), out_bf16=torch.empty(M, D, **bf))),
        ("q2   bias->bf16", (M, D, D), dict(a=x1b, b=wgt(D, D), bias=bias(D), out_bf16=torch.empty(M, D, **bf))),
        ("o2   bias+res->f32", (M, D, D), dict(a=a2, b=wgt(D, D), bias=bias(D), res=x1, out_f32=torch.empty(M, D, **f32))),
        ("ff1  bias+gelu->bf16", (M, 4 * D, D), dict(a=h2, b=wgt(4 * D, D), bias=bias(4 * D), act=ops.ACT_GELU_TANH, out_bf16=torch.empty(M, 4 * D, **bf))),
        ("ff2  bias+gate+res->f32", (M, D, 4 * D), dict(a=g, b=wgt(D, 4 * D), bias=bias(D), gate=mod[:, 5 * D:6 * D], rows_per_gate=N, res=x2,
                                                     out_f32=torch.empty(M, D, **f32))),
    ]

    def once():
        for _, _, kw in calls:
            kw = dict(kw)
            ops.gemm(kw.pop("a"), kw.pop("b"), **kw)

    flops = sum(2.0 * m * n * k for _, (m, n, k), _ in calls)
    for _ in range(10):
        once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        once()
    e1.record()
    torch.cuda.synchronize()
    iters = max(20, int(min_ms / (e0.elapsed_time(e1) / 20)))
    sampler = ClockSampler(device_index, period=0.01).start()
    e0.record()
    for _ in range(iters):
        once()
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / iters
    ach = flops / (ms * 1e-3) / 1e12
    # per-shape times (separate short loops; they explain the aggregate, the aggregate is the roofline number)
    per_shape = {}
    for name, (m, n, k), kw in calls:
        kw = dict(kw)
        a_, b_ = kw.pop("a"), kw.pop("b")
        e0.record()
        for _ in range(50):
            ops.gemm(a_, b_, **kw)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 50 * 1e3
        per_shape[name] = {"us": round(us, 2), "tflops": round(2.0 * m * n * k / us / 1e6, 1)}
    traffic = None
    tp = ROOT / "profiles" / "r02_gemm_traffic.json"  # written by scripts/ncu_traffic.py from an `ncu --set full` capture
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text())
        except Exception:
            traffic = None
    return {"bound": "tensor",
            "kernel": "gemm2_bf16_kernel (CTA-pair tcgen05/TMEM/TMA GEMM, streaming TMA epilogue; half of the iteration): the 6 "
                      "GEMMs of one DiT block's forward at M=%d with their in-iteration epilogues" % M,
            "achieved": ach, "peak": peaks["burst"], "unit": "TFLOP/s", "frac": ach / peaks["burst"],
            "frac_of_sustained_peak": ach / peaks["sustained"], "traffic": traffic,
            "peak_source": peaks["source"] + ", burst figure (kernel timed alone; the loop runs %.0f ms, i.e. at sustained clocks)" % (ms * iters),
            "launches_timed": iters * len(calls), "avg_launch_us": ms * 1e3 / len(calls), "clocks": clocks,
            "per_shape": per_shape,
            "l2": "six GEMMs with distinct operands per round, ~430 MB touched per round, above the 126 MB L2"}


def library_baseline(config_name, B, lora_r, steps=2):
    """BASELINE.md section 4 'GPU library baseline': the SAME iteration as stock PyTorch would run it on this GPU today
    -- the oracle modules moved to the device under bf16 autocast (cuBLASLt GEMMs, fused SDPA attention, eager
    elementwise, autograd, torch.optim.AdamW).  The closest runnable stand-in for the reference's diffusers + xformers
    path; explanatory, outside every timed region of the product, never part of `value`."""
    from oracle import dit as oracle_dit
    from oracle.iteration import RepairedIteration
    from tdm_b200.harness import synthetic_text
    dev = torch.device("cuda", torch.cuda.current_device())
    oracle_dit.Attention.use_sdpa = True
    try:
        cfg, student, fake, teacher, sch = build_oracle(config_name, lora_r)
        for m in (student, fake, teacher):
            m.to(dev)
        ctx, mask, uctx, umask = (t.to(dev) for t in synthetic_text(B, seed=0, Cc=cfg.caption_channels))
        ctx, uctx = ctx.to(torch.bfloat16), uctx.to(torch.bfloat16)
        it = RepairedIteration(student, fake, teacher, sch, uctx, umask, lowp=torch.bfloat16)
        it.alphas, it.sigmas = it.alphas.to(dev), it.sigmas.to(dev)
        it.predictor.alphas, it.predictor.sigmas = it.alphas, it.sigmas
        sch.alphas_cumprod = sch.alphas_cumprod.to(dev)
        S = cfg.sample_size
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            for i in range(2 + steps):
                if i == 2:
                    torch.cuda.synchronize()
                    e0.record()
                torch.manual_seed(100 + i)
                noise = torch.randn(B, 4, S, S).to(torch.bfloat16).to(dev)
                out = it.step(noise, ctx, mask)
            e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        res = {"value": 1e3 / ms, "unit": UNIT, "ms_per_step": ms, "steps": steps, "warmup": 2,
               "what": "oracle iteration on cuda:0, torch.autocast(bf16): cuBLASLt + SDPA + eager elementwise + autograd + "
                       "torch.optim.AdamW (fp32 master weights), same batch and passes; timed outside the product's regions",
               "loss": float(out["loss"])}
    finally:
        oracle_dit.Attention.use_sdpa = False
        del_names = ("student", "fake", "teacher", "it")
        for n in del_names:
            locals().pop(n, None)
        torch.cuda.empty_cache()
    return res


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
                         cfg=4.5, total_steps=900, use_huber=True, use_separate=True, lowp=lowp, grad_sync=gs)
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
        fl = 4 * f["fwd"] * Bs  # merged weights: a student forward costs what a teacher forward costs
        sampling = {"metric": "images_per_s_4nfe", "value": world * Bs / (ms_s * 1e-3), "e2e": world * Bs / (ms_se * 1e-3),
                    "unit": "final latents/s (4 NFE, timesteps 999/856/665/399), summed over GPUs", "batch_per_gpu": Bs,
                    "ms_per_batch": ms_s, "tflops_per_gpu": fl / (ms_s * 1e-3) / 1e12,
                    "frac_of_sustained_peak": fl / (ms_s * 1e-3) / 1e12 / peaks["sustained"]}
        # ---- the step after sampling: TAESD decode to RGB, training.py:113-117 ----
        try:
            from tdm_b200.taesd import AutoencoderTiny, decode_to_images
            vae = AutoencoderTiny(device=dev, seed=3)
            for _ in range(2):
                img = decode_to_images(vae, sample_4nfe(student, sched, snoise_d, sctx_d, smask_d, lowp))
            barrier()
            d0, d1, d2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            d0.record()
            for _ in range(reps):
                lat = sample_4nfe(student, sched, snoise_d, sctx_d, smask_d, lowp)
            d1.record()
            for _ in range(reps):
                img = decode_to_images(vae, lat)
            d2.record()
            barrier()
            td = torch.tensor([d0.elapsed_time(d1) / reps, d1.elapsed_time(d2) / reps], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(td, op=dist.ReduceOp.MAX)
            ms_lat, ms_dec = td.tolist()
            Hs = S * 8
            conv_flops = 2.0 * 64 * 64 * 9 * (9 * S * S + 10 * 4 * S * S + 10 * 16 * S * S + 4 * 64 * S * S) + 2.0 * 9 * (4 * 64 * S * S + 64 * 3 * 64 * S * S)
            sampling["rgb"] = {"metric": "rgb_images_per_s_4nfe_plus_taesd", "value": world * Bs / ((ms_lat + ms_dec) * 1e-3),
                               "unit": f"{Hs}x{Hs} RGB images/s (4-NFE student + TAESD decode, clamp to [0,1]), summed over GPUs",
                               "decode_ms_per_batch": ms_dec, "decode_images_per_s_per_gpu": Bs / (ms_dec * 1e-3),
                               "decode_tflops_per_gpu": conv_flops * Bs / (ms_dec * 1e-3) / 1e12,
                               "decode_share_of_sampling": ms_dec / (ms_lat + ms_dec),
                               "image_range": [float(img.min()), float(img.max())]}
            del vae, img
        except Exception as e:  # the decoder is a 'next' row: its failure must not take the headline down
            sampling["rgb"] = {"value": None, "failed": repr(e)[:300]}

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
    roof = gemm_roofline(cfg, B, peaks, local)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": workload_string(args.config, cfg, B, args.lora_rank, world),
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
            sec, cores, _ = cpu_iteration_seconds(args.config, 1, args.lora_rank, 1, 0)
            v = 1.0 / (sec * B)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"bounded sample: one oracle iteration (plain PyTorch fp32, SDPA attention, {cores} host threads) at "
                                              f"batch 1 of {B}, full {args.config} model, no warm-up: {sec:.1f} s; batch-{B} rate = 1/({B} x t). "
                                              f"`bench.py --impl reference` times the full batch when host RAM allows"}
        except Exception as e:  # the baseline is reported, never required for the GPU numbers
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": f"failed: {e!r}"}
    if world == 1 and not args.no_library_baseline:
        try:
            del trainer, student, fake, teacher
            torch.cuda.empty_cache()
            line["gpu_library_baseline"] = library_baseline(args.config, B, args.lora_rank)
            line["gpu_library_baseline"]["speedup_of_this_repo"] = value / line["gpu_library_baseline"]["value"]
        except Exception as e:  # explanatory only
            line["gpu_library_baseline"] = {"value": None, "failed": repr(e)[:300]}
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
    ap.add_argument("--no-library-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
