#!/bin/bash
# BASELINE.json configs[1..4] at N GPUs of one box (under `gpurun --gpus N`): bench lines -> gpurun_out/bench_<tag>_n<N>_*.json
# usage: scripts/multi_gpu_configs.sh <tag> <N>
set -u
TAG=${1:-r02}; N=${2:-8}
OUT=gpurun_out; mkdir -p $OUT
PORT=29600
run() { name=$1; shift; PORT=$((PORT + 1))
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N "$@" \
    > $OUT/bench_${TAG}_n${N}_$name.json 2> $OUT/bench_${TAG}_n${N}_$name.err || { echo "$name FAILED"; tail -5 $OUT/bench_${TAG}_n${N}_$name.err; }
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_${TAG}_n${N}_$name.json").read().strip().splitlines()[-1])
    print("$name", "n_gpus", d["n_gpus"], round(d["value"], 1), d["unit"], round(d["ms_per_step"], 3), "ms/step", "e2e", (d.get("e2e") or {}).get("value"))
except Exception as e:
    print("$name: no line", e)
PY
}
run c2 --steps 20 --warmup 5 --sustained 0
run c3 --dim-neck 32 --freq 32 --batch 128 --len-crop 256 --steps 20 --warmup 5 --sustained 0
run c4 --n-bins 513 --steps 20 --warmup 5 --sustained 0
run c5_convert --workload convert --utterances 4096 --steps 2 --warmup 3
run frontend --workload frontend --utterances 4096 --steps 5 --warmup 3
