"""Repaired training driver with the reference's CLI (src/main.py:73-594 with the defects of SURVEY.md 3.5 repaired as
DESIGN.md section 2 lists them), running entirely on libtdm_b200.

    python -m tdm_b200.main --train_batch_size 8 --max_train_steps 100 --use_huber --use_separate --cfg 4.5 ...
    python -m tdm_b200.main --run_demo /path/to/train_tdm_demo.py <flags>   # executes the reference's unmodified entry
                                                                            # script with this driver installed as src.main

The reference's setup needs the network (T5, TAESD, the PixArt checkpoint, JourneyDB), so the offline
harness provides random-init models and synthetic T5 embeddings (north_star); `Transformer2DModel._pretrained_provider`
lets a deployment plug real weights in.  Flags are the reference's (src/args.py);
its parser is used verbatim when TDM_REFERENCE_ROOT names a checkout, otherwise the TDM-relevant subset below (unknown
flags are accepted and ignored)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch


DEFAULT_MAX_STEPS = 100  # the synthetic harness has no dataset length to derive num_train_epochs * steps from


def _reference_parser():
    """The reference's own `parse_args` (src/args.py imports without diffusers / accelerate), loaded by file path when a
    checkout is named by TDM_REFERENCE_ROOT; None otherwise.  Nothing outside that variable is searched."""
    root = os.environ.get("TDM_REFERENCE_ROOT")
    if not root:
        return None
    path = os.path.join(root, "src", "args.py")
    if not os.path.exists(path):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("tdm_reference_args", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.parse_args


IGNORED_FLAGS = ("gradient_accumulation_steps", "scale_lr", "num_train_epochs", "use_8bit_adam")


def parse_args(argv=None):
    ref_parse = _reference_parser()
    if ref_parse is not None:
        old = sys.argv
        sys.argv = [old[0]] + list(argv if argv is not None else old[1:])
        try:
            return ref_parse()
        finally:
            sys.argv = old
    p = argparse.ArgumentParser(description="TDM distillation (B200-native)")
    p.add_argument("--pretrained_model_name_or_path", default="PixArt-alpha/PixArt-XL-2-512x512")
    p.add_argument("--output_dir", default="tdm-output")
    p.add_argument("--seed", type=int, default=None)
    p.add_argument("--train_batch_size", type=int, default=16)
    p.add_argument("--max_train_steps", type=int, default=None)
    p.add_argument("--num_train_epochs", type=int, default=100)
    p.add_argument("--learning_rate", type=float, default=1e-4)
    p.add_argument("--adam_beta1", type=float, default=0.9)
    p.add_argument("--adam_beta2", type=float, default=0.999)
    p.add_argument("--adam_weight_decay", type=float, default=1e-2)
    p.add_argument("--adam_epsilon", type=float, default=1e-8)
    p.add_argument("--max_grad_norm", type=float, default=1.0)
    p.add_argument("--lr_scheduler", default="constant")
    p.add_argument("--lr_warmup_steps", type=int, default=500)
    p.add_argument("--mixed_precision", default=None, choices=["no", "fp16", "bf16"])
    p.add_argument("--checkpointing_steps", type=int, default=500)
    p.add_argument("--checkpoints_total_limit", type=int, default=None)
    p.add_argument("--resume_from_checkpoint", default=None)
    p.add_argument("--use_ema", action="store_true")
    p.add_argument("--gradient_checkpointing", action="store_true")
    p.add_argument("--enable_xformers_memory_efficient_attention", action="store_true")
    p.add_argument("--dataset_name", default=None)
    p.add_argument("--train_data_dir", default=None)
    p.add_argument("--local_rank", type=int, default=-1)
    p.add_argument("--prediction_type", default=None)
    p.add_argument("--cfg", type=float, default=4.5)
    p.add_argument("--total_steps", type=int, default=900)
    p.add_argument("--use_huber", action="store_true")
    p.add_argument("--use_separate", action="store_true")
    p.add_argument("--use_reg", action="store_true")
    p.add_argument("--gradient_accumulation_steps", type=int, default=1)
    p.add_argument("--scale_lr", action="store_true")
    p.add_argument("--use_8bit_adam", action="store_true")
    args, _unknown = p.parse_known_args(argv)
    return args


def lr_at(args, step, total=None):
    """diffusers.optimization.get_scheduler for the schedules the README uses (constant / constant_with_warmup /
    cosine_with_restarts with one cycle == cosine), evaluated on the host per step."""
    import math
    name, warm = args.lr_scheduler, args.lr_warmup_steps
    total = max(1, total if total is not None else (args.max_train_steps or DEFAULT_MAX_STEPS))
    if name == "constant":
        return args.learning_rate
    if step < warm:
        return args.learning_rate * float(step) / float(max(1, warm))
    if name == "constant_with_warmup":
        return args.learning_rate
    prog = float(step - warm) / float(max(1, total - warm))
    if name in ("cosine", "cosine_with_restarts"):
        return args.learning_rate * max(0.0, 0.5 * (1.0 + math.cos(math.pi * (prog % 1.0 if name == "cosine_with_restarts" and prog < 1.0 else min(prog, 1.0)))))
    if name == "linear":
        return args.learning_rate * max(0.0, 1.0 - prog)
    raise ValueError(f"unsupported lr_scheduler {name}")


def main(argv=None):
    from . import checkpoint as ck
    from . import dp, harness
    from .training import EMAModel, TdmTrainer, draw_iteration_noise, sample_4nfe

    args = parse_args(argv)
    config = os.environ.get("TDM_B200_CONFIG", "pixart-512")  # "tiny" for smoke runs (BASELINE.json configs[0])
    rank, local, world = dp.init_from_env("nccl")
    dev = torch.device("cuda", local)
    if args.seed is not None:
        torch.manual_seed(args.seed)  # accelerate.utils.set_seed seeds every rank identically (main.py:117-118)
    for flag in IGNORED_FLAGS:  # accepted for CLI compatibility, without effect here
        v, d = getattr(args, flag, None), {"gradient_accumulation_steps": 1, "num_train_epochs": 100}.get(flag, False)
        if v not in (None, d) and rank == 0:
            print(f"tdm_b200.main: --{flag}={v} is accepted but has no effect in this driver", file=sys.stderr)
    lowp = torch.float16 if args.mixed_precision == "fp16" else torch.bfloat16
    cfg = harness.CONFIGS[config]
    S = cfg["sample_size"]
    B = args.train_batch_size
    student, fake, teacher = harness.make_models(config, seed=args.seed or 0, device=dev, lora_rank=64)
    dp.broadcast_params([student, fake, teacher])
    if args.gradient_checkpointing:
        student.enable_gradient_checkpointing()
    student.enable_xformers_memory_efficient_attention()
    ctx, mask, uctx, umask = harness.synthetic_text(B, seed=(args.seed or 0) + 17 + rank, Cc=cfg["caption_channels"])
    ctx, mask = ctx.to(dev).to(torch.bfloat16), mask.to(dev)
    sched = harness.make_scheduler()
    if args.prediction_type is not None:
        sched.register_to_config(prediction_type=args.prediction_type)  # main.py:447-449
    gs = dp.GradSync() if world > 1 else None
    trainer = TdmTrainer(student, fake, teacher, sched, uctx.to(dev).to(torch.bfloat16), umask.to(dev), cfg=args.cfg,
                         total_steps=args.total_steps, use_huber=args.use_huber, use_separate=args.use_separate, lowp=lowp,
                         lr=args.learning_rate, lr_fake=args.learning_rate, max_grad_norm=args.max_grad_norm,
                         weight_decay=args.adam_weight_decay, grad_sync=gs)
    for opt in (trainer.opt_student, trainer.opt_fake):
        opt.betas, opt.eps = (args.adam_beta1, args.adam_beta2), args.adam_epsilon
    ema = EMAModel(student) if args.use_ema else None
    os.makedirs(args.output_dir, exist_ok=True)
    global_step = ck.load_checkpoint(args.output_dir, args.resume_from_checkpoint, trainer, ema) \
        if args.resume_from_checkpoint else 0
    max_steps = args.max_train_steps or DEFAULT_MAX_STEPS
    fixed_noise = draw_iteration_noise(4, (S, S), lowp, dev)  # main.py:544-547 (fixed validation noise)
    vctx, vmask, _, _ = harness.synthetic_text(4, seed=(args.seed or 0) + 911, Cc=cfg["caption_channels"])  # 4 fixed prompts
    vctx, vmask = vctx.to(dev).to(torch.bfloat16), vmask.to(dev)
    # data stream: the model init above is identical on every rank; from here each rank draws its own noise / indices
    # (accelerate's set_seed would leave all ranks on one stream, i.e. N copies of the same batch)
    if world > 1 and not args.resume_from_checkpoint:
        torch.manual_seed((args.seed or 0) + 1000 + rank)
        torch.cuda.manual_seed((args.seed or 0) + 1000 + rank)
    log = open(os.path.join(args.output_dir, f"train_log_rank{rank}.jsonl"), "a")
    t0 = time.time()
    while global_step < max_steps:
        lr = lr_at(args, global_step, max_steps)
        trainer.opt_student.lr = trainer.opt_fake.lr = lr
        noise = draw_iteration_noise(B, (S, S), lowp, dev)                     # main.py:443
        out = trainer.step(noise, ctx, mask)
        if ema is not None:
            ema.step(student.parameters())                                       # main.py:553-554
        global_step += 1
        if global_step % 10 == 0 or global_step == max_steps:                    # one host sync per 10 steps (main.py:531-532 every step)
            rec = {"step": global_step, "loss_instruct": float(out["loss"]), "loss_fake": float(out["loss_fake"]), "lr": lr,
                   "it_per_s": global_step / max(1e-9, time.time() - t0)}
            log.write(json.dumps(rec) + "\n")
            log.flush()
            if rank == 0:
                print(json.dumps(rec), flush=True)
        if rank == 0 and global_step % 50 == 0:                                  # main.py:542-549, training.py:107-112
            lat4 = sample_4nfe(student, sched, fixed_noise, vctx, vmask, lowp)
            torch.save(lat4.cpu(), os.path.join(args.output_dir, f"fixed_4step_{global_step}.pt"))
        if rank == 0 and global_step % args.checkpointing_steps == 0:            # main.py:563-587
            ck.save_checkpoint(args.output_dir, global_step, trainer, args.checkpoints_total_limit, ema=ema)
    log.close()
    return 0


def run_demo(script, argv):
    """Execute the reference's unmodified entry script (train_tdm_demo.py: `from src.main import main; main()`) with
    this driver installed as `src.main`."""
    import runpy
    import types
    src = types.ModuleType("src")
    src.__path__ = []  # a package
    shim = types.ModuleType("src.main")
    shim.main = lambda: main(argv)
    sys.modules["src"], sys.modules["src.main"] = src, shim
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--run_demo":
        run_demo(sys.argv[2], sys.argv[3:])
    else:
        sys.exit(main())
