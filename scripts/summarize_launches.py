"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel (shares of the captured window)."""
import collections
import csv
import re
import sys

src, dst = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else None)
rows = list(csv.DictReader(l for l in open(src) if l.startswith('"')))
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for r in rows:
    name = re.sub(r"\(.*", "", re.sub(r"^void ", "", r["Kernel Name"]))
    t = float(r["Metric Value"].replace(",", "")) / 1e3
    agg[name][0] += 1
    agg[name][1] += t
    tot += t
out = [f"# {len(rows)} launches, {tot:.0f} us total (ncu: serialised, cold cache -- compare shares)", "kernel,launches,total_us,share,avg_us"]
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{k},{n},{t:.1f},{t / tot:.4f},{t / n:.2f}")
txt = "\n".join(out) + "\n"
if dst:
    open(dst, "w").write(txt)
print("\n".join(out[:24]))
