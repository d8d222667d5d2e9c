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


def trainer_state(trainer):
    """Everything needed to continue a run: fp32 master arenas, optimizer moments + step counts, RNG states."""
    def opt_state(opt):
        return {"t": opt.t, "moments": [(m.detach().cpu(), v.detach().cpu()) for m, v in opt.state]}
    st = {
        "student": trainer.student._flat.detach().cpu(),
        "student_lora": None if trainer.student._lflat is None else trainer.student._lflat.detach().cpu(),
        "fake": trainer.fake._flat.detach().cpu(),
        "opt_student": opt_state(trainer.opt_student),
        "opt_fake": opt_state(trainer.opt_fake),
        "rng_cpu": torch.get_rng_state(),
        "rng_cuda": torch.cuda.get_rng_state() if torch.cuda.is_available() else None,
    }
    return st


def load_trainer_state(trainer, st):
    with torch.no_grad():
        trainer.student._flat.copy_(st["student"])
        if st["student_lora"] is not None:
            trainer.student._lflat.copy_(st["student_lora"])
        trainer.fake._flat.copy_(st["fake"])
        for opt, key in ((trainer.opt_student, "opt_student"), (trainer.opt_fake, "opt_fake")):
            opt.t = st[key]["t"]
            for (m, v), (ms, vs) in zip(opt.state, st[key]["moments"]):
                m.copy_(ms)
                v.copy_(vs)
    for mdl in (trainer.student, trainer.fake):
        mdl._flat16_version = mdl._lflat16_version = -1  # bf16 copies follow the restored masters
        mdl.mark_updated()
    torch.set_rng_state(st["rng_cpu"])
    if st.get("rng_cuda") is not None and torch.cuda.is_available():
        torch.cuda.set_rng_state(st["rng_cuda"])


def save_checkpoint(output_dir, global_step, trainer, checkpoints_total_limit=None, lora_prefix="lora_unet"):
    """main.py:563-587 + the kohya LoRA file users load into a pipeline."""
    os.makedirs(output_dir, exist_ok=True)
    rotate_checkpoints(output_dir, checkpoints_total_limit)
    save_path = os.path.join(output_dir, f"checkpoint-{global_step}")
    os.makedirs(save_path, exist_ok=True)
    torch.save(trainer_state(trainer), os.path.join(save_path, "trainer_state.pt"))
    if trainer.student._lflat is not None:
        from .utils import get_module_kohya_state_dict
        torch.save(get_module_kohya_state_dict(trainer.student, lora_prefix, torch.float16),
                   os.path.join(save_path, "lora_kohya.pt"))
    return save_path


def load_checkpoint(output_dir, resume_from_checkpoint, trainer):
    path, step = resolve_resume(output_dir, resume_from_checkpoint)
    if path is None:
        return 0
    st = torch.load(os.path.join(output_dir, path, "trainer_state.pt"), map_location="cpu", weights_only=False)
    load_trainer_state(trainer, st)
    return step
