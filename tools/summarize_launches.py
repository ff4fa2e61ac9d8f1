"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(row["Metric Unit"], v / 1e3)
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
print(f"| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k[:90]}` | {n} | {t:.1f} | {100 * t / tot:.1f}% | {t / n:.1f} |")
print(f"\ntotal {tot / 1e3:.2f} ms over {sum(n for n, _ in agg.values())} launches (cold-cache, serialised: compare shares)")
