#!/bin/bash
# Runs on the GPU box (under gpurun): plain bench, then the ncu launch list of one steady-state step
# of the SAME command, then one `--set full` capture of the dominant kernel.
# usage: scripts/gpu_profile.sh <tag> <kernel-regex> [extra bench args...]
set -u
TAG=${1:-r01}; KREGEX=${2:-gemm_nt_taps}; shift 2 || true
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline $*"
$CMD > $OUT/bench_${TAG}_plain.json 2> $OUT/bench_${TAG}_plain.err || { echo "plain run failed"; tail -20 $OUT/bench_${TAG}_plain.err; exit 1; }
LPS=$(python -c "import json,sys; print(json.loads(open('$OUT/bench_${TAG}_plain.json').read().strip().splitlines()[-1])['gpu_launches_per_step'])")
echo "launches per step: $LPS"
SKIP=$((LPS * 3 + 200))
# ncu cannot replay the cluster+cooperative K-split BPTT kernel (it records nan and aborts): exclude it by name; its
# device time comes from the CUDA-event family timing in the bench JSON instead
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(?!.*lstm_tc_bwd_ks).*$' -s $SKIP -c $LPS --csv --log-file $OUT/launches_${TAG}.csv $CMD > $OUT/ncu_${TAG}_launches.log 2>&1
echo "ncu launch list rc=$?"
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 40 -c 3 -f -o $OUT/prof_${TAG} $CMD > $OUT/ncu_${TAG}_full.log 2>&1
echo "ncu full rc=$?"
