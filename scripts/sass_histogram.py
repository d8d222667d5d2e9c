"""SASS opcode evidence for the shipped library: per kernel, how many tcgen05 / TMEM / TMA instructions it holds
(`cuobjdump -sass`; mnemonics per /opt/skills/guides/B200_PROFILING.md: UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,
UTMALDG/UTMASTG = TMA tensor load/store, UTCBAR = tcgen05.commit, UTCCP = tcgen05.cp).
    python scripts/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parents[1] / "tdm_b200" / "libtdm_b200.so"
KEYS = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "UTCCP", "SYNCS", "HMMA", "MUFU.EX2",
        "REDG", "RED.", "ATOMG", "LDG", "STG", "LDS", "STS", "BAR.SYNC"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            per[cur]["_total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    per[cur][k] += 1
    total = collections.Counter()
    print(f"# {LIB.name}: SASS opcode counts per kernel (sm_100a); columns: " + " ".join(KEYS))
    for name, c in per.items():
        total.update(c)
        if any(c[k] for k in KEYS[:10]):
            print(f"{name[:110]:110s} total={c['_total']:6d} " + " ".join(f"{k}={c[k]}" for k in KEYS if c[k]))
    print("# library totals: " + " ".join(f"{k}={total[k]}" for k in KEYS if total[k]) + f" kernels={len(per)}")


if __name__ == "__main__":
    sys.exit(main())
