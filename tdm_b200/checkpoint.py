"""Checkpoint / resume of a TDM run, mirroring what the reference gets from accelerate.save_state / load_state
(main.py:379-404 resume resolution incl. "latest"; main.py:563-587 rotation with --checkpoints_total_limit) and the
kohya-format LoRA export (utils.py:76-93).  Host-side plumbing: tensors are the models' flat arenas and the fused
optimizer's flat moment buffers, so a checkpoint is a handful of large contiguous tensors."""
from __future__ import annotations

import os
import shutil

import torch


def _ckpt_dirs(output_dir):
    dirs = [d for d in os.listdir(output_dir) if d.startswith("checkpoint")]
    return sorted(dirs, key=lambda x: int(x.split("-")[1]))


def rotate_checkpoints(output_dir, checkpoints_total_limit):
    """main.py:565-582: before saving a new checkpoint keep at most `limit - 1` existing ones. Returns removed names."""
    if checkpoints_total_limit is None:
        return []
    checkpoints = _ckpt_dirs(output_dir)
    removed = []
    if len(checkpoints) >= checkpoints_total_limit:
        num_to_remove = len(checkpoints) - checkpoints_total_limit + 1
        removed = checkpoints[0:num_to_remove]
        for name in removed:
            shutil.rmtree(os.path.join(output_dir, name))
    return removed


def resolve_resume(output_dir, resume_from_checkpoint):
    """main.py:379-404: returns (checkpoint dir name or None, global_step)."""
    if not resume_from_checkpoint:
        return None, 0
    if resume_from_checkpoint != "latest":
        path = os.path.basename(resume_from_checkpoint)
    else:
        dirs = _ckpt_dirs(output_dir) if os.path.isdir(output_dir) else []
        path = dirs[-1] if len(dirs) > 0 else None
    if path is None:
        return None, 0
    return path, int(path.split("-")[1])


def kohya_lora_state_dict(adapters, prefix, dtype, lora_alpha):
    """kohya-ss LoRA naming (what utils.py:76-93 produces and `pipe.load_lora_weights` consumes, README.md:49, 96, 119),
    built from the adapter name table: `<path.to.linear>.lora_A.weight` -> `<prefix>_<path_to_linear>.lora_down.weight`,
    lora_B -> lora_up, plus one `<...>.alpha` entry right after each down matrix."""
    role = {"lora_A": "lora_down", "lora_B": "lora_up"}
    alpha = torch.tensor(lora_alpha).to(dtype)
    out = {}
    for key, w in adapters.items():
        path = key.split("base_model.model.", 1)[-1]
        linear, ab, leaf = path.rsplit(".", 2)
        stem = "_".join([prefix] + linear.split("."))
        out[f"{stem}.{role[ab]}.{leaf}"] = w.to(dtype)
        if ab == "lora_A":
            out[f"{stem}.alpha"] = alpha.clone()
    return out


def _join_streams(trainer):
    """A deferred fake-score optimiser step (TdmTrainer.defer_fake_opt) may still be running on the trainer's side
    stream: make the current stream wait for it before the arenas are read or overwritten."""
    trainer.sync()
    side = getattr(trainer, "_side", None)
    if side is not None:
        torch.cuda.current_stream().wait_stream(side)


def trainer_state(trainer, ema=None):
    """Everything needed to continue a run, as plain tensors / ints (loadable with weights_only=True): fp32 master
    arenas, optimizer moments + step counts, EMA shadow + its step counter (--use_ema), RNG states."""
    _join_streams(trainer)
    st = {
        "student": trainer.student._flat.detach().cpu(),
        "fake": trainer.fake._flat.detach().cpu(),
        "rng_cpu": torch.get_rng_state(),
    }
    if trainer.student._lflat is not None:
        st["student_lora"] = trainer.student._lflat.detach().cpu()
    if torch.cuda.is_available():
        st["rng_cuda"] = torch.cuda.get_rng_state(trainer.device)
    for key, opt in (("opt_student", trainer.opt_student), ("opt_fake", trainer.opt_fake)):
        st[key + ".t"] = int(opt.t)
        for i, (m, v) in enumerate(opt.state):
            st[f"{key}.m{i}"], st[f"{key}.v{i}"] = m.detach().cpu(), v.detach().cpu()
    if ema is not None:
        st["ema.step"] = int(ema.optimization_step)
        for i, sh in enumerate(ema.shadow):
            st[f"ema.shadow{i}"] = sh.detach().cpu()
    return st


def load_trainer_state(trainer, st, ema=None):
    _join_streams(trainer)
    with torch.no_grad():
        trainer.student._flat.copy_(st["student"])
        if "student_lora" in st:
            trainer.student._lflat.copy_(st["student_lora"])
        trainer.fake._flat.copy_(st["fake"])
        for key, opt in (("opt_student", trainer.opt_student), ("opt_fake", trainer.opt_fake)):
            opt.t = int(st[key + ".t"])
            for i, (m, v) in enumerate(opt.state):
                m.copy_(st[f"{key}.m{i}"])
                v.copy_(st[f"{key}.v{i}"])
        if ema is not None and "ema.step" in st:
            ema.optimization_step = int(st["ema.step"])
            for i, sh in enumerate(ema.shadow):
                sh.copy_(st[f"ema.shadow{i}"])
    for mdl in (trainer.student, trainer.fake):
        mdl._flat16_version = mdl._lflat16_version = -1  # bf16 copies follow the restored masters
        mdl.mark_updated()
    torch.set_rng_state(st["rng_cpu"])
    if "rng_cuda" in st and torch.cuda.is_available():
        torch.cuda.set_rng_state(st["rng_cuda"], trainer.device)


def save_checkpoint(output_dir, global_step, trainer, checkpoints_total_limit=None, lora_prefix="lora_unet", ema=None):
    """main.py:563-587 + the kohya LoRA file users load into a pipeline."""
    os.makedirs(output_dir, exist_ok=True)
    rotate_checkpoints(output_dir, checkpoints_total_limit)
    save_path = os.path.join(output_dir, f"checkpoint-{global_step}")
    os.makedirs(save_path, exist_ok=True)
    torch.save(trainer_state(trainer, ema), os.path.join(save_path, "trainer_state.pt"))
    if trainer.student._lflat is not None:
        from .utils import get_module_kohya_state_dict
        torch.save(get_module_kohya_state_dict(trainer.student, lora_prefix, torch.float16),
                   os.path.join(save_path, "lora_kohya.pt"))
    return save_path


def load_checkpoint(output_dir, resume_from_checkpoint, trainer, ema=None):
    """main.py:379-404: a checkpoint that does not exist is reported and the run starts fresh (returns 0)."""
    path, step = resolve_resume(output_dir, resume_from_checkpoint)
    file = None if path is None else os.path.join(output_dir, path, "trainer_state.pt")
    if file is None or not os.path.exists(file):
        print(f"Checkpoint '{resume_from_checkpoint}' does not exist. Starting a new training run.")  # main.py:390-393
        return 0
    st = torch.load(file, map_location="cpu", weights_only=True)
    load_trainer_state(trainer, st, ema)
    return step
