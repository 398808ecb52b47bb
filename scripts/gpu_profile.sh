#!/bin/bash
# Runs on the GPU box (under gpurun): the plain bench, then -- only if that exited 0 -- the ncu launch list of one
# steady-state step of the SAME command (one ncu invocation per gpurun call).
# usage: scripts/gpu_profile.sh <tag> [extra bench args...]
set -u
TAG=${1:-r01}; shift 1 || true
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline $*"
$CMD > $OUT/bench_${TAG}_plain.json 2> $OUT/bench_${TAG}_plain.err || { echo "plain run failed"; tail -20 $OUT/bench_${TAG}_plain.err; exit 1; }
LPS=$(python -c "import json,sys; print(json.loads(open('$OUT/bench_${TAG}_plain.json').read().strip().splitlines()[-1])['gpu_launches_per_step'])")
echo "launches per step: $LPS"
SKIP=$((LPS * 3 + 200))
# The persistent recurrence kernels (lstm_tc_*) spin on counters published by the other CTAs of a cluster / cooperative
# launch; ncu's kernel replay records nan for them and aborts, so they are excluded by name.  Their device time comes from
# the CUDA-event family timing in the bench JSON (roofline.families) instead.  The window is several steps long (the -k
# filter also counts torch's own kernels); summarize_profiles.py cuts exactly one step out of it between two L1-loss kernels.
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:^(?!.*lstm_tc).*$' -s $SKIP -c $((LPS * 4)) --csv --log-file $OUT/launches_${TAG}.csv $CMD > $OUT/ncu_${TAG}_launches.log 2>&1
echo "ncu launch list rc=$?"
