#!/usr/bin/env python
"""Turn gpurun_out/ profiler outputs into small tracked summaries under profiles/.

  python scripts/summarize_profiles.py <tag>

Reads gpurun_out/launches_<tag>.csv (ncu --metrics gpu__time_duration.sum launch list of ONE
steady-state step), gpurun_out/prof_<tag>.ncu-rep (ncu --set full of the dominant kernel) and
gpurun_out/bench_<tag>*.json, writes profiles/<tag>_launches.md, profiles/<tag>_ncu_full.csv and
copies the bench JSON line."""
import csv
import io
import json
import os
import re
import subprocess
import sys
from collections import OrderedDict

tag = sys.argv[1]
G, P = "gpurun_out", "profiles"
os.makedirs(P, exist_ok=True)


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void ", "").replace("avc::", "")


rows = []
path = f"{G}/launches_{tag}.csv"
if os.path.exists(path):
    lines = [l for l in open(path) if l.startswith('"')]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((short(r["Kernel Name"]), r["Grid Size"], r["Block Size"], float(r["Metric Value"])))
    # exactly one step: from one L1-loss forward kernel (once per step) to the next
    marks = [i for i, r in enumerate(rows) if r[0].startswith("loss_fwd_kernel<1>") or r[0].startswith("loss_fwd_kernel<true>")]
    if len(marks) >= 2:
        rows = rows[marks[0]:marks[1]]
    agg = OrderedDict()
    for k, g, b, ns in rows:
        a = agg.setdefault(k, [0, 0.0, set()])
        a[0] += 1
        a[1] += ns
        a[2].add(g)
    tot = sum(a[1] for a in agg.values()) or 1.0
    with open(f"{P}/{tag}_launches.md", "w") as f:
        f.write(f"# {tag}: ncu launch list of one steady-state step (gpu__time_duration.sum, --clock-control none)\n\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write(f"{len(rows)} launches, {tot/1e6:.2f} ms total kernel time\n\n| kernel | launches | total ms | share | grids |\n|---|---:|---:|---:|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            grids = ", ".join(sorted(a[2])[:4])
            f.write(f"| `{k}` | {a[0]} | {a[1]/1e6:.3f} | {100*a[1]/tot:.1f}% | {grids} |\n")
    print(open(f"{P}/{tag}_launches.md").read())

rep = f"{G}/prof_{tag}.ncu-rep"
if os.path.exists(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    lines = [l for l in out.splitlines() if l.startswith('"')]
    rd = list(csv.reader(io.StringIO("\n".join(lines))))
    hdr, units, data = rd[0], rd[1], rd[2:]
    keep = re.compile(r"Kernel Name|Grid Size|Block Size|gpu__time_duration.sum|dram__bytes_(read|write)\.sum$|dram__throughput.avg.pct|"
                      r"sm__pipe_tensor.*cycles_active.avg.pct|sm__inst_executed_pipe_tensor|sm__warps_active.avg.pct|launch__registers_per_thread|"
                      r"launch__occupancy_limit|sm__throughput.avg.pct|l1tex__data_bank_conflicts_pipe_lsu.sum$|smsp__inst_executed.sum$|"
                      r"sm__pipe_fma_cycles_active.avg.pct|sm__inst_executed_pipe_fma|lts__t_sector_hit_rate.pct|sm__cycles_elapsed.max|launch__shared_mem_per_block")
    idx = [i for i, h in enumerate(hdr) if keep.search(h)]
    with open(f"{P}/{tag}_ncu_full.csv", "w") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for d in data:
            w.writerow([short(d[i]) if hdr[i] == "Kernel Name" else d[i] for i in idx])
    print(open(f"{P}/{tag}_ncu_full.csv").read()[:3000])

for fn in os.listdir(G):
    if fn.startswith(f"bench_{tag}") and fn.endswith(".json") and os.path.getsize(f"{G}/{fn}") > 0:
        open(f"{P}/{fn}", "w").write(open(f"{G}/{fn}").read())
