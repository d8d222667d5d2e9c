"""Summarise an `ncu --page source --csv` dump: SASS opcode mix by executed count and the top stall sites."""
import collections
import csv
import sys

lines = open(sys.argv[1]).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(lines[start:]))
ops = collections.Counter()
samples = collections.Counter()
tot_inst = tot_samp = 0
for r in rows:
    src = r["Source"].strip()
    op = src.split()[0] if src else "?"
    if op.startswith("@"):
        op = src.split()[1]
    n = int(float(r["Instructions Executed"] or 0))
    s = int(float(r["# Samples"] or 0))
    ops[op.split(".")[0]] += n
    samples[op.split(".")[0]] += s
    tot_inst += n
    tot_samp += s
print(f"warp instructions executed: {tot_inst}, stall samples: {tot_samp}")
print("opcode, executed, share, samples_share")
for op, n in ops.most_common(22):
    print(f"{op},{n},{n / max(tot_inst, 1):.3f},{samples[op] / max(tot_samp, 1):.3f}")
top = sorted(rows, key=lambda r: -int(float(r["# Samples"] or 0)))[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]
stall_cols = [c for c in rows[0].keys() if c and c.startswith("stall_")]
print("top stall sites:")
for r in top:
    st = sorted(((int(float(r[c] or 0)), c) for c in stall_cols), reverse=True)[:2]
    print(r["Address"][-5:], r["# Samples"], r["Source"][:70], st)
