"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of the step).
usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.md"""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(row["Metric Unit"], v)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"# ncu launch list summary: {path}")
print(f"\n{sum(v[0] for v in agg.values())} launches, {tot/1e3:.2f} ms of kernel time (cold-cache, serialised: compare SHARES)\n")
print("| share | launches | avg us | kernel |\n|---:|---:|---:|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {v[1]/tot*100:.2f}% | {v[0]} | {v[1]/v[0]:.1f} | `{k}` |")
# template instances of one kernel (epilogue flavours of the conv kernel) summed
fam = collections.defaultdict(lambda: [0, 0.0])
for k, v in agg.items():
    base = re.sub(r"<.*", "", k).replace("void ", "")
    fam[base][0] += v[0]
    fam[base][1] += v[1]
print("\nPer kernel family (all template instances):\n\n| share | launches | avg us | kernel family |\n|---:|---:|---:|---|")
for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1])[:8]:
    print(f"| {v[1]/tot*100:.2f}% | {v[0]} | {v[1]/v[0]:.1f} | `{k}` |")
