#!/usr/bin/env python
"""App-range replay counters of the persistent recurrence launches -> profiles/<tag>_lstm_counters.md + <tag>_traffic.json.

  python scripts/summarize_lstm_counters.py <tag>

Reads gpurun_out/ncu_<tag>_lstm_apprange_{fwd,bwd}_h{1024,512}.csv (scripts/profile_r02_lstm.sh) and the free-running launch
times of the same shapes from gpurun_out/bench_lstm_<tag>.log (scripts/bench_lstm.py, half-mode entry points)."""
import csv
import io
import json
import os
import sys

tag = sys.argv[1]
G, P = "gpurun_out", "profiles"
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "": 1.0, "inst": 1.0, "%": 1.0, "cycle": 1.0,
        "sector": 1.0}

free = {}
path = f"{G}/bench_lstm_{tag}.log"
if os.path.exists(path):
    for line in open(path):
        if line.startswith("{"):
            d = json.loads(line)
            if d["kernel"].endswith("_h") and d["B"] == 256 and d["T"] == 128:
                free[("fwd" if "fwd" in d["kernel"] else "bwd", d["H"])] = d["ms"]

rows, traffic = [], {}
for H in (1024, 512):
    for which in ("fwd", "bwd"):
        path = f"{G}/ncu_{tag}_lstm_apprange_{which}_h{H}.csv"
        if not os.path.exists(path):
            continue
        m = {}
        lines = [l for l in open(path) if l.startswith('"')]
        for r in csv.DictReader(io.StringIO("".join(lines))):
            m[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * UNIT.get(r.get("Metric Unit", ""), 1.0)
        if not m:
            continue
        ms = free.get((which, H))
        dram = m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        xbar = m.get("l1tex__m_xbar2l1tex_read_bytes.sum", 0.0)
        traffic[f"{which}_h{H}"] = {"dram_bytes": dram, "l2_to_sm_bytes": xbar, "lts_bytes": m.get("lts__t_bytes.sum", 0.0)}
        rows.append((which, H, ms, m, dram, xbar))

os.makedirs(P, exist_ok=True)
with open(f"{P}/{tag}_lstm_counters.md", "w") as f:
    f.write(f"# {tag}: hardware counters of the weight-stationary recurrence kernels (`ncu --replay-mode app-range`)\n\n"
            "One launch of `lstm_tc_fwd_ws_kernel` / `lstm_tc_bwd_ws_kernel` at B=256, T=128 between cudaProfilerStart/Stop\n"
            "(`scripts/ncu_lstm_range.py`, `scripts/profile_r02_lstm.sh`; why app-range: `profiles/r02_lstm_counters.md`).  Bandwidths divide\n"
            "the counted BYTES by the un-profiled launch time of `scripts/bench_lstm.py` (half-mode entry points, same shapes).\n\n"
            "| launch | time (ms, free-running) | us/step | DRAM read MB | DRAM write MB | L2 bytes (lts__t_bytes) GB | L2->SM (xbar2l1tex) GB | "
            "L2->SM TB/s | per SM GB/s | shared-memory wavefronts (LSU) M | tensor instr | tensor pipe active % (in-profiler) | range time in profiler ms |\n"
            "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
    for which, H, ms, m, dram, xbar in rows:
        ncta = 128
        tbs = xbar / (ms * 1e-3) / 1e12 if ms else float("nan")
        f.write(f"| {which} H={H} | {ms if ms is None else round(ms, 3)} | {'' if ms is None else round(ms * 1e3 / 128, 2)} | "
                f"{m.get('dram__bytes_read.sum', 0) / 1e6:.0f} | {m.get('dram__bytes_write.sum', 0) / 1e6:.0f} | {m.get('lts__t_bytes.sum', 0) / 1e9:.2f} | "
                f"{xbar / 1e9:.2f} | {tbs:.2f} | {tbs * 1e3 / ncta:.0f} | {m.get('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 0) / 1e6:.1f} | "
                f"{m.get('sm__inst_executed_pipe_tensor.sum', 0):.0f} | {m.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 0):.1f} | "
                f"{m.get('gpu__time_duration.sum', 0) / 1e6:.3f} |\n")
json.dump({"source": f"ncu --replay-mode app-range, one launch at B=256 T=128 (profiles/{tag}_lstm_counters.md)", "per_launch": traffic},
          open(f"{P}/{tag}_traffic.json", "w"), indent=1)
print(open(f"{P}/{tag}_lstm_counters.md").read())
