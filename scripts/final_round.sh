#!/bin/bash
# Round-end measurement suite on ONE B200 (run under gpurun): tests, smoke, bench lines of every config and arm.
# Multi-GPU lines: scripts/multi_gpu_configs.sh / multi_gpu_extra.sh; profiler evidence: scripts/profile_r02.sh, profile_r02_lstm.sh.
# usage: scripts/final_round.sh <tag>
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -2
run() { name=$1; shift; python bench.py "$@" > $OUT/bench_${TAG}_$name.json 2> $OUT/bench_${TAG}_$name.err || { echo "$name FAILED"; tail -5 $OUT/bench_${TAG}_$name.err; }; python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_${TAG}_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"],1), d["unit"], round(d["ms_per_step"],3), "ms", (d.get("e2e") or {}).get("value"))
except Exception as e: print("$name: no line", e)
PY
}
run half --steps 50 --warmup 10
run reference --impl reference --steps 2 --warmup 1
run torch_gpu --impl torch-gpu --steps 10 --warmup 3
run fp32 --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline --no-modes --sustained 0
run half_c3 --dim-neck 32 --freq 32 --batch 128 --len-crop 256 --steps 20 --warmup 5 --no-cpu-baseline --no-modes --sustained 0
run half_c4 --n-bins 513 --steps 20 --warmup 5 --no-cpu-baseline --no-modes --sustained 0
run convert --workload convert --steps 2 --warmup 3
run frontend --workload frontend --steps 5 --warmup 3
run wav --workload wav --steps 10 --warmup 3
run loader --workload loader
run dvector --workload dvector
