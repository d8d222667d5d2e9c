"""In-tree build of libtdm_b200.so (sm_100a only) with plain nvcc.

`python -m tdm_b200.build` or `__graft_entry__.build()`.  Objects are rebuilt only when the source
(or a header) is newer.  The .so is git-ignored but travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
OBJ = ROOT / "build"
LIB = ROOT / "libtdm_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def sources():
    return sorted(CSRC.glob("*.cu"))


def _needs(obj: Path, src: Path, headers) -> bool:
    if not obj.exists():
        return True
    t = obj.stat().st_mtime
    return any(p.stat().st_mtime > t for p in [src, *headers])


def build(verbose: bool = False, force: bool = False) -> Path:
    nvcc = os.environ.get("NVCC", "nvcc")
    OBJ.mkdir(exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list((ROOT.parent / "include").glob("*.h"))
    jobs = []
    for src in sources():
        obj = OBJ / (src.stem + ".o")
        if force or _needs(obj, src, headers):
            cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if verbose:
            sys.stderr.write(r.stderr)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    objs = [str(OBJ / (s.stem + ".o")) for s in sources()]
    if jobs or not LIB.exists():
        # libcuda's tensor-map encoder is resolved through cudaGetDriverEntryPoint at run time; the gradient exchange
        # is torch.distributed (tdm_b200/dp.py), so the library links neither NCCL nor libcuda.
        cmd = [nvcc, "-shared", "-o", str(LIB), *objs, "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(p)
