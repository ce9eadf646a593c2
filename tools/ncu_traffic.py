"""Summarise an `ncu --set full` report (exported with `ncu -i X.ncu-rep --page raw --csv > X.csv`) per kernel: launches, mean duration, DRAM bytes
(dram__bytes_read.sum + dram__bytes_write.sum), tensor-pipe and memory-throughput percentages.  Writes a markdown table and, with --json, the
profiles/r2_traffic.json file bench.py reads its `roofline.traffic` from.

    python tools/ncu_traffic.py gpurun_out/r2_full.csv --md profiles/r2_b_ncu_full_summary.md --json profiles/r2_traffic.json --batch 16 --precision f16x3
"""
import argparse
import collections
import csv
import json
import re

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--md")
ap.add_argument("--json")
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--precision", default="f16x3")
a = ap.parse_args()

rows = list(csv.reader(open(a.csv)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[hdr]
col = {n: i for i, n in enumerate(H)}
want = {"dur": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
        "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor2": "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
        "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l2_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed", "occ": "sm__warps_active.avg.pct_of_peak_sustained_active"}
units = rows[hdr + 1]


def num(r, key):
    i = col.get(want[key])
    if i is None or i >= len(r) or r[i] in ("", "n/a"):
        return None
    v = float(r[i].replace(",", ""))
    u = units[i].lower()
    if key == "dur":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(u, 1e-3)
    if key in ("rd", "wr"):
        v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    return v


agg = collections.OrderedDict()
for r in rows[hdr + 2:]:
    if len(r) <= col["Kernel Name"]:
        continue
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("dsb::", "").replace("<unnamed>::", "")
    g = agg.setdefault(name, collections.defaultdict(list))
    for k in want:
        v = num(r, k)
        if v is not None:
            g[k].append(v)
mean = lambda x: sum(x) / len(x) if x else None
lines = ["| kernel | launches | mean us | DRAM MB / launch (rd + wr) | DRAM GB/s | tensor pipe % | DRAM % of peak | L2 % | SM % |", "|---|---|---|---|---|---|---|---|---|"]
out = {}
for name, g in sorted(agg.items(), key=lambda kv: -sum(kv[1]["dur"])):
    n = len(g["dur"])
    d, rd, wr = mean(g["dur"]), mean(g["rd"]), mean(g["wr"])
    tp = mean(g["tensor"]) if g["tensor"] else mean(g["tensor2"])
    byt = (rd or 0) + (wr or 0)
    fmt = lambda v, f="%.1f": "-" if v is None else f % v
    lines.append(f"| `{name}` | {n} | {d:.1f} | {byt / 1e6:.2f} ({(rd or 0) / 1e6:.2f} + {(wr or 0) / 1e6:.2f}) | {byt / d / 1e3:.0f} | {fmt(tp)} | {fmt(mean(g['dram_pct']))} | "
                 f"{fmt(mean(g['l2_pct']))} | {fmt(mean(g['sm_pct']))} |")
    out[name] = {"launches": n, "mean_us": d, "dram_bytes_per_launch": byt, "tensor_pipe_pct": tp, "dram_pct": mean(g["dram_pct"])}
print("\n".join(lines))
if a.md:
    open(a.md, "a").write("\n".join(lines) + "\n")
if a.json:
    try:
        j = json.load(open(a.json))
    except Exception:
        j = {}
    gem = (next((v for k, v in out.items() if "gemm_f16x3_pair" in k), None) or next((v for k, v in out.items() if "gemm_tcgen05_pair" in k), None)
           or next((v for k, v in out.items() if "gemm_tcgen05" in k), None))
    if gem:
        j[f"gemm_{a.precision}_B{a.batch}"] = {"dram_bytes_per_launch": gem["dram_bytes_per_launch"], "mean_us_under_ncu": gem["mean_us"], "launches": gem["launches"],
                                                 "tensor_pipe_pct": gem["tensor_pipe_pct"],
                                                 "source": f"ncu --set full capture of tools/profile_kernels.py ({a.csv}), mean over the denoiser-layer GEMM launches"}
    j.setdefault("kernels", {})[a.precision] = out
    json.dump(j, open(a.json, "w"), indent=1)
