"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one diffusion step, per-kernel totals and shares.

    python tools/summarize_launches.py gpurun_out/launches.csv [step_index] > profiles/<name>.md
"""
import collections
import csv
import re
import sys

path = sys.argv[1]
step = int(sys.argv[2]) if len(sys.argv) > 2 else 0
with open(path) as f:
    rows = list(csv.DictReader([l for l in f if not l.startswith("==")]))
names = [r["Kernel Name"] for r in rows]
marks = [i for i, n in enumerate(names) if "embed_tokens" in n]
start = marks[step]
end = marks[step + 1] if step + 1 < len(marks) else len(rows)
agg, tot = collections.OrderedDict(), 0.0
for r in rows[start:end]:
    n = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")[:70]
    v = float(r["Metric Value"].replace(",", ""))
    v = v / 1000 if r["Metric Unit"] == "ns" else (v * 1000 if r["Metric Unit"] == "ms" else v)
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
print(f"# ncu launch list, diffusion step {step}: {end - start} launches, {tot:.1f} us of kernel time (cold-cache, serialised: compare shares)\n")
print("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{n}` | {c} | {v:.1f} | {100 * v / tot:.1f}% | {v / c:.1f} |")
