#!/usr/bin/env python
"""Turn the ncu artefacts a gpurun call brought back (gpurun_out/) into the small, tracked summaries under profiles/.
usage: python scripts/summarize_profiles.py r1"""
import collections
import csv
import json
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out = []


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
        agg.setdefault(name, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    res = ["| kernel | launches | total us | mean us | share |", "|---|---:|---:|---:|---:|"]
    for k, v in agg.items():
        res.append("| `%s` | %d | %.1f | %.1f | %.3f |" % (k[:70], len(v), sum(v), sum(v) / len(v), sum(v) / tot))
    return res, tot


def raw(rep, keys):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (units[i], vals[i]) for i, h in enumerate(hdr)}
    return ["| %s | %s | %s |" % (k, d[k][1], d[k][0]) for k in keys if k in d]


KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_dim_x",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]

out.append("# ncu summaries, round %s\n" % tag)
lp = "gpurun_out/launches_%s.csv" % tag
if os.path.exists(lp):
    res, tot = launches(lp)
    out.append("## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, bench.py --steps 2 --warmup 1)\n")
    out.append("Per-launch times are cold-cache and serialised: compare SHARES. Total %.1f us over all captured launches "
               "(setup + 3 device-resident steps + 3 end-to-end steps).\n" % tot)
    out += res
for name in ("recurrent", "gemm"):
    rep = "gpurun_out/prof_%s_%s.ncu-rep" % (name, tag)
    if os.path.exists(rep):
        out.append("\n## `ncu --set full` : %s\n" % name)
        out.append("| metric | value | unit |\n|---|---|---|")
        out += raw(rep, KEYS)
os.makedirs("profiles", exist_ok=True)
open("profiles/%s_ncu_summary.md" % tag, "w").write("\n".join(out) + "\n")
for f in ("bench_ours.json", "bench_ref.json"):
    p = os.path.join("gpurun_out", f)
    if os.path.exists(p):
        open("profiles/%s_%s" % (tag, f), "w").write(open(p).read())
print("\n".join(out))
