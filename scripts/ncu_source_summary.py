"""Summarise an `ncu --page source --csv` dump (possibly several kernel sections): per section the SASS opcode mix by
executed count and the top stall sites.  usage: ncu_source_summary.py dump.csv [n_top] [section_index]"""
import collections
import csv
import sys

lines = open(sys.argv[1]).read().splitlines()
n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
starts = [i for i, l in enumerate(lines) if l.startswith('"Address"')]
names = [lines[i - 1] for i in starts]
ends = [s - 1 for s in starts[1:]] + [len(lines)]
start, end = starts[which], ends[which]
print("section", which, "of", len(starts), names[which][:120])
rows = list(csv.DictReader(lines[start:end]))
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
top = sorted(rows, key=lambda r: -int(float(r["# Samples"] or 0)))[:n_top]
stall_cols = [c for c in rows[0].keys() if c and c.startswith("stall_") and "Not Issued" not in c]
print("top stall sites:")
for r in top:
    st = sorted(((int(float(r[c] or 0)), c) for c in stall_cols), reverse=True)[:2]
    print(r["Address"][-5:], r["# Samples"], r["Instructions Executed"], r["Source"][:70], st)
