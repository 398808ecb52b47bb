#!/bin/bash
# Round-end measurement suite on ONE B200 (run under gpurun): tests, smoke, bench lines of every config, launch list, front-end ncu.
# usage: scripts/final_round.sh <tag>
set -u
TAG=${1:-r01e}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -2
run() { name=$1; shift; python bench.py "$@" > $OUT/bench_${TAG}_$name.json 2> $OUT/bench_${TAG}_$name.err || { echo "$name FAILED"; tail -5 $OUT/bench_${TAG}_$name.err; }; python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_${TAG}_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"],1), d["unit"], round(d["ms_per_step"],3), "ms", (d.get("e2e") or {}).get("value"))
except Exception as e: print("$name: no line", e)
PY
}
run half --steps 20 --warmup 5
run reference --impl reference --steps 2 --warmup 1
run tf32 --precision tf32 --steps 5 --warmup 3 --no-cpu-baseline
run fp32 --precision fp32 --steps 3 --warmup 3 --no-cpu-baseline
run half_c3 --dim-neck 32 --freq 32 --batch 128 --len-crop 256 --steps 10 --warmup 3 --no-cpu-baseline
run half_c4 --n-bins 513 --steps 10 --warmup 3 --no-cpu-baseline
run convert --workload convert --steps 3 --warmup 3
run frontend --workload frontend --steps 5 --warmup 3
bash scripts/gpu_profile.sh $TAG 2>&1 | tail -3
ncu --set full --clock-control none --import-source on -k regex:fe_ -s 5 -c 3 -o $OUT/fe_${TAG} -f python bench.py --workload frontend --utterances 1024 --steps 1 --warmup 1 > $OUT/fe_${TAG}_ncu.log 2>&1
ncu -i $OUT/fe_${TAG}.ncu-rep --page raw --csv > $OUT/fe_${TAG}_raw.csv 2>/dev/null
ls -la $OUT/fe_${TAG}.ncu-rep $OUT/fe_${TAG}_raw.csv
