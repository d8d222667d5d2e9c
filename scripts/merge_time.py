"""Time of one LoRA merge of the full student (280 adapted linears, 3.6 GB of algorithmic traffic)."""
import json, sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "scripts")
from _timing import timeit
from tdm_b200 import harness
student, _, _ = harness.make_models("pixart-512", seed=0, device="cuda", lora_rank=64, teacher=False, fake=False)
def run(i):
    student._eff_key = None
    student._refresh_merged()
us, mhz = timeit(run, iters=10)
print(json.dumps({"case": "lora merge, full student", "us": round(us, 1), "sm_mhz": mhz, "GBps_alg": round(3.6e9 / us / 1e3, 1)}))
