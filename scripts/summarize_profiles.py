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
    dram = {}       # launch ID -> DRAM bytes (read + write), when the launch list was taken with the dram__bytes metrics
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r.get("Metric Unit", "byte"), 1.0)
            dram[r["ID"]] = dram.get(r["ID"], 0.0) + float(r["Metric Value"].replace(",", "")) * scale
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            val = float(r["Metric Value"].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r.get("Metric Unit", "ns"), 1.0)
            rows.append((short(r["Kernel Name"]), r["Grid Size"], r["Block Size"], val, dram.get(r["ID"])))
    # exactly one step: from one L1-loss forward kernel (once per step) to the next
    marks = [i for i, r in enumerate(rows) if r[0].startswith("loss_fwd_kernel<1>") or r[0].startswith("loss_fwd_kernel<true>")]
    if len(marks) >= 2:
        rows = rows[marks[0]:marks[1]]
    agg = OrderedDict()
    for k, g, b, ns, db in rows:
        a = agg.setdefault(k, [0, 0.0, set(), 0.0])
        a[0] += 1
        a[1] += ns
        a[2].add(g)
        a[3] += db or 0.0
    tot = sum(a[1] for a in agg.values()) or 1.0
    with open(f"{P}/{tag}_launches.md", "w") as f:
        f.write(f"# {tag}: ncu launch list of one steady-state step (gpu__time_duration.sum, --clock-control none)\n\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write(f"{len(rows)} launches, {tot/1e6:.2f} ms total kernel time\n\n| kernel | launches | total ms | share | DRAM MB (read + write, all launches) | grids |\n|---|---:|---:|---:|---:|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            grids = ", ".join(sorted(a[2])[:4])
            f.write(f"| `{k}` | {a[0]} | {a[1]/1e6:.3f} | {100*a[1]/tot:.1f}% | {a[3]/1e6:.1f} | {grids} |\n")
    if dram:
        # DRAM bytes per C-ABI call of the GEMM / BatchNorm families (bench.py: roofline.traffic).  The split reductions of the
        # weight-gradient GEMMs (wgrad_reduce_*) serve both TN families: shared out by launch count.
        def tot_of(pred):
            return sum(a[3] for k, a in agg.items() if pred(k)), sum(a[0] for k, a in agg.items() if pred(k))
        tn_h, n_tn_h = tot_of(lambda k: k.startswith("tc_gemm2_tn_kernel<2>") or k.startswith("tc_gemm_kernel<1, 2"))
        tn_f, n_tn_f = tot_of(lambda k: k.startswith("tc_gemm2_tn_kernel<4>") or k.startswith("tc_gemm_kernel<1, 4"))
        red, n_red = tot_of(lambda k: k.startswith("wgrad_reduce"))
        nt_h, n_nt_h = tot_of(lambda k: k.startswith("tc_gemm2_nt_kernel<2>") or k.startswith("tc_gemm_kernel<0, 2"))
        fam = {}
        if n_tn_h:
            fam["avc_gemm_tn_taps_h"] = (tn_h + red * n_tn_h / max(1, n_tn_h + n_tn_f)) / n_tn_h
        if n_tn_f:
            fam["avc_gemm_tn_taps"] = (tn_f + red * n_tn_f / max(1, n_tn_h + n_tn_f)) / n_tn_f
        if n_nt_h:
            fam["avc_gemm_nt_taps_hw"] = nt_h / n_nt_h
        for fname, kern in (("avc_bn_act_bwd_reduce_y", "bn_bwd_reduce_cols_kernel"), ("avc_bn_act_bwd_apply_y", "bn_bwd_apply_cols_kernel"),
                            ("avc_bn_act_fwd_h", "bn_act_fwd_cols_kernel")):
            b, n = tot_of(lambda k: k.startswith(kern))
            if n:
                fam[fname] = b / n
        json.dump({"source": f"ncu launch list of one bench step with dram__bytes_read/write.sum (gpurun_out/launches_{tag}.csv, scripts/gpu_profile.sh); "
                             "bytes per C-ABI call = kernel launches of the family summed / calls", "per_call": fam},
                  open(f"{P}/{tag}_family_traffic.json", "w"), indent=1)
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
