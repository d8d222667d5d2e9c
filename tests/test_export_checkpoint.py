"""kohya LoRA export against the REAL reference function (utils.py:76-93, run unmodified with a stub `peft` that
returns the same adapter dict), and the checkpoint rotation / resume resolution of main.py:379-404, 563-587. CPU."""
import os
import sys
import types

import pytest
import torch

from pathlib import Path

from oracle import reference_import

ROOT = Path(__file__).resolve().parents[1]


class _FakeLoraModel:
    """Duck type of what get_module_kohya_state_dict touches: peft_state_dict() and peft_config."""

    def __init__(self):
        g = torch.Generator().manual_seed(0)
        self.sd = {}
        for blk in range(2):
            for lin in ("attn1.to_q", "attn2.to_out.0", "ff.net.0.proj", "ff.net.2"):
                self.sd[f"base_model.model.transformer_blocks.{blk}.{lin}.lora_A.weight"] = torch.randn(4, 8, generator=g)
                self.sd[f"base_model.model.transformer_blocks.{blk}.{lin}.lora_B.weight"] = torch.randn(8, 4, generator=g)
        self.peft_config = {"default": types.SimpleNamespace(lora_alpha=32.0, r=4)}

    def peft_state_dict(self, adapter_name="default"):
        return dict(self.sd)


def test_kohya_names_known_answer():
    from tdm_b200.utils import get_module_kohya_state_dict
    out = get_module_kohya_state_dict(_FakeLoraModel(), "lora_unet", torch.float16)
    assert "lora_unet_transformer_blocks_0_attn1_to_q.lora_down.weight" in out
    assert "lora_unet_transformer_blocks_1_ff_net_0_proj.lora_up.weight" in out
    assert "lora_unet_transformer_blocks_0_attn2_to_out_0.alpha" in out
    assert out["lora_unet_transformer_blocks_0_attn1_to_q.alpha"].item() == 32.0
    assert all(v.dtype == torch.float16 for v in out.values())
    assert len(out) == 2 * 4 * 3  # down + up + alpha per adapted linear


@pytest.mark.skipif(not reference_import.available(), reason="/root/reference not on this machine")
def test_kohya_export_equals_reference_function():
    from tdm_b200.utils import get_module_kohya_state_dict
    m = _FakeLoraModel()
    stub = types.ModuleType("peft")
    stub.get_peft_model_state_dict = lambda module, adapter_name="default": module.peft_state_dict(adapter_name)
    old = sys.modules.get("peft")
    sys.modules["peft"] = stub
    try:
        for k in [k for k in sys.modules if k.startswith("tdm_reference_src")]:
            del sys.modules[k]
        utils, _, _ = reference_import.load()
        utils.get_peft_model_state_dict = stub.get_peft_model_state_dict
        ref = utils.get_module_kohya_state_dict(m, "lora_unet", torch.float16)
    finally:
        if old is not None:
            sys.modules["peft"] = old
        else:
            sys.modules.pop("peft", None)
    ours = get_module_kohya_state_dict(m, "lora_unet", torch.float16)
    assert list(ref.keys()) == list(ours.keys())
    assert all(torch.equal(ref[k], ours[k]) for k in ref)


def test_rotation_and_latest(tmp_path):
    from tdm_b200 import checkpoint as ck
    out = str(tmp_path)
    for step in (100, 200, 1000, 300):
        os.makedirs(os.path.join(out, f"checkpoint-{step}"))
    assert ck.resolve_resume(out, "latest") == ("checkpoint-1000", 1000)      # numeric, not lexicographic, order
    assert ck.resolve_resume(out, "/some/where/checkpoint-200") == ("checkpoint-200", 200)
    assert ck.resolve_resume(out, None) == (None, 0)
    removed = ck.rotate_checkpoints(out, 3)  # keep at most limit-1 = 2 before the new one is written
    assert removed == ["checkpoint-100", "checkpoint-200"]
    assert sorted(os.listdir(out)) == ["checkpoint-1000", "checkpoint-300"]
    assert ck.rotate_checkpoints(out, None) == []
    empty = str(tmp_path / "none")
    os.makedirs(empty)
    assert ck.resolve_resume(empty, "latest") == (None, 0)


@pytest.mark.gpu
def test_checkpoint_round_trip_on_gpu(tmp_path):
    from tdm_b200 import checkpoint as ck, harness
    from tdm_b200.training import TdmTrainer, draw_iteration_noise
    from tdm_b200.utils import get_module_kohya_state_dict
    torch.manual_seed(0)
    student, fake, teacher = harness.make_models("tiny", seed=0, device="cuda", lora_rank=64)
    ctx, mask, uctx, umask = harness.synthetic_text(2, L=24, seed=0)
    tr = TdmTrainer(student, fake, teacher, harness.make_scheduler(), uctx.cuda(), umask.cuda(), lr=1e-3, lr_fake=1e-3)
    for _ in range(2):
        tr.step(draw_iteration_noise(2, (32, 32), torch.bfloat16, "cuda"), ctx.cuda(), mask.cuda())
    path = ck.save_checkpoint(str(tmp_path), 2, tr, checkpoints_total_limit=2)
    assert os.path.exists(os.path.join(path, "lora_kohya.pt"))
    want = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in ck.trainer_state(tr).items() if k in ("student_lora", "fake")}
    tr.step(draw_iteration_noise(2, (32, 32), torch.bfloat16, "cuda"), ctx.cuda(), mask.cuda())  # move away
    assert not torch.equal(tr.student._lflat.cpu(), want["student_lora"])
    assert ck.load_checkpoint(str(tmp_path), "latest", tr) == 2
    assert torch.equal(tr.student._lflat.cpu(), want["student_lora"]) and torch.equal(tr.fake._flat.cpu(), want["fake"])
    assert tr.opt_fake.t == 2
    koh = get_module_kohya_state_dict(student, "lora_unet", torch.float16)
    assert len(koh) == 2 * 10 * 3 and any(k.endswith("to_q.lora_down.weight") for k in koh)
    out = tr.step(draw_iteration_noise(2, (32, 32), torch.bfloat16, "cuda"), ctx.cuda(), mask.cuda())
    assert torch.isfinite(out["loss"]) and torch.isfinite(out["loss_fake"])


@pytest.mark.gpu
def test_ema_matches_diffusers_formula():
    from tdm_b200 import harness
    from tdm_b200.training import EMAModel
    student, _, _ = harness.make_models("tiny", seed=0, device="cuda", lora_rank=64, teacher=False, fake=False)
    ema = EMAModel(student, decay=0.9999)
    ref = [p.clone() for p in student.flat_params()]
    for step in range(1, 4):
        with torch.no_grad():
            for p in student.flat_params():
                p.add_(0.01 * step)
        ema.step(student.parameters())
        d = 0.0 if step == 1 else min(0.9999, (1 + (step - 1)) / (10 + (step - 1)))
        ref = [r - (1 - d) * (r - p) for r, p in zip(ref, student.flat_params())]
    for r, s in zip(ref, ema.shadow):
        assert torch.allclose(r, s, rtol=1e-6, atol=1e-7)


@pytest.mark.gpu
def test_driver_runs_with_the_reference_cli(tmp_path, monkeypatch):
    """The repaired driver with the reference's flags (README.md:158-176) on BASELINE.json configs[0] (tiny DiT)."""
    import json
    from tdm_b200 import main as drv
    monkeypatch.setenv("TDM_B200_CONFIG", "tiny")
    out = str(tmp_path / "run")
    argv = ["--train_batch_size", "2", "--max_train_steps", "12", "--use_huber", "--use_separate", "--cfg", "4.5",
            "--learning_rate", "2e-5", "--lr_scheduler", "cosine_with_restarts", "--lr_warmup_steps", "5",
            "--checkpointing_steps", "10", "--checkpoints_total_limit", "2", "--output_dir", out, "--seed", "3",
            "--dataset_name", "synthetic", "--mixed_precision", "bf16", "--use_ema"]
    assert drv.main(argv) == 0
    recs = [json.loads(l) for l in open(os.path.join(out, "train_log_rank0.jsonl"))]
    assert recs[-1]["step"] == 12 and all(r["loss_instruct"] == r["loss_instruct"] for r in recs)  # finite, logged
    assert os.path.isdir(os.path.join(out, "checkpoint-10"))
    # resume picks the latest checkpoint up and continues to the new limit
    assert drv.main(argv[:3] + ["14"] + argv[4:] + ["--resume_from_checkpoint", "latest"]) == 0
    recs = [json.loads(l) for l in open(os.path.join(out, "train_log_rank0.jsonl"))]
    assert recs[-1]["step"] == 14


@pytest.mark.gpu
def test_checkpoint_with_deferred_optimizer_and_ema(tmp_path):
    """with defer_fake_opt the fake-score AdamW of the last step may still be running on the side stream when
    a checkpoint is taken; trainer_state()/load_trainer_state() join it.  The EMA shadow and its step counter travel
    with the checkpoint; the file holds plain tensors / ints only (weights_only=True)."""
    from tdm_b200 import checkpoint as ck, harness
    from tdm_b200.training import EMAModel, TdmTrainer, draw_iteration_noise
    torch.manual_seed(0)
    student, fake, teacher = harness.make_models("tiny", seed=0, device="cuda", lora_rank=64)
    ctx, mask, uctx, umask = harness.synthetic_text(2, L=24, seed=0)
    tr = TdmTrainer(student, fake, teacher, harness.make_scheduler(), uctx.cuda(), umask.cuda(), lr=1e-3, lr_fake=1e-3,
                    defer_fake_opt=True)
    assert tr.defer_fake_opt
    ema = EMAModel(student)
    for _ in range(3):
        tr.step(draw_iteration_noise(2, (32, 32), torch.bfloat16, "cuda"), ctx.cuda(), mask.cuda())
        ema.step(student.parameters())
    st = ck.trainer_state(tr, ema)           # no explicit sync by the caller
    torch.cuda.synchronize()
    assert torch.equal(st["fake"], fake._flat.cpu())                       # not a torn copy of a step in flight
    assert torch.equal(st["opt_fake.m0"], tr.opt_fake.state[0][0].cpu())
    path = ck.save_checkpoint(str(tmp_path), 3, tr, ema=ema)
    raw = torch.load(os.path.join(path, "trainer_state.pt"), weights_only=True)
    assert raw["ema.step"] == 3 and raw["opt_fake.t"] == 3
    shadow = [s.clone() for s in ema.shadow]
    tr.step(draw_iteration_noise(2, (32, 32), torch.bfloat16, "cuda"), ctx.cuda(), mask.cuda())
    ema.step(student.parameters())
    assert ck.load_checkpoint(str(tmp_path), "latest", tr, ema) == 3
    assert ema.optimization_step == 3 and all(torch.equal(a, b) for a, b in zip(shadow, ema.shadow))
    assert torch.equal(fake._flat.cpu(), st["fake"]) and tr.opt_fake.t == 3
    # main.py:389-393: a named checkpoint that is not there -> message, fresh run
    assert ck.load_checkpoint(str(tmp_path), "checkpoint-77", tr, ema) == 0
    assert ck.load_checkpoint(str(tmp_path / "nowhere"), "latest", tr, ema) == 0


@pytest.mark.gpu
def test_run_demo_executes_the_reference_entry_script(tmp_path, monkeypatch):
    """`python -m tdm_b200.main --run_demo train_tdm_demo.py <flags>`: the reference's entry script (train_tdm_demo.py:
    21-24, `from src.main import main` + `main()` under the __main__ guard) runs unchanged with this driver installed as
    src.main.  The script body is restated here (the reference tree does not travel to the GPU box)."""
    import json
    import subprocess
    script = tmp_path / "train_tdm_demo.py"
    script.write_text('"""entry point"""\n\nfrom src.main import main\n\nif __name__ == "__main__":\n    main()\n')
    out = str(tmp_path / "demo")
    env = dict(os.environ, TDM_B200_CONFIG="tiny", PYTHONPATH=str(ROOT))
    env.pop("TDM_REFERENCE_ROOT", None)
    r = subprocess.run([sys.executable, "-m", "tdm_b200.main", "--run_demo", str(script), "--train_batch_size", "2",
                        "--max_train_steps", "10", "--use_huber", "--use_separate", "--cfg", "4.5", "--output_dir", out,
                        "--seed", "1", "--dataset_name", "synthetic", "--mixed_precision", "bf16",
                        "--lr_scheduler", "cosine", "--lr_warmup_steps", "2", "--gradient_accumulation_steps", "2"],
                       env=env, capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    assert "--gradient_accumulation_steps=2 is accepted but has no effect" in r.stderr
    recs = [json.loads(l) for l in open(os.path.join(out, "train_log_rank0.jsonl"))]
    assert recs[-1]["step"] == 10 and recs[-1]["lr"] > 0  # cosine over the resolved 10 steps, not collapsed to 0


def test_lr_schedule_uses_the_resolved_step_count():
    from tdm_b200 import main as drv
    a = drv.parse_args(["--dataset_name", "x", "--lr_scheduler", "cosine", "--lr_warmup_steps", "10", "--learning_rate", "1e-4"])
    assert a.max_train_steps is None
    vals = [drv.lr_at(a, s, drv.DEFAULT_MAX_STEPS) for s in (0, 5, 10, 55, 99)]
    assert vals[0] == 0.0 and abs(vals[1] - 5e-5) < 1e-12 and abs(vals[2] - 1e-4) < 1e-12
    assert abs(vals[3] - 5e-5) < 1e-9 and 0 < vals[4] < 1e-6
    assert drv.lr_at(a, 50) == drv.lr_at(a, 50, drv.DEFAULT_MAX_STEPS)   # same default inside and outside the loop
