#!/bin/bash
# Round-2 (third session) profiler evidence after the weight-stationary recurrences: launch list of one bench step with DRAM
# bytes per launch, app-range replay counters of the four recurrence launches, free-running per-step latencies and traces.
set -u
OUT=gpurun_out; mkdir -p $OUT
bash scripts/gpu_profile.sh r02c 2>&1 | tail -3
TAG=r02c bash scripts/profile_r02_lstm.sh 2>&1 | tail -8
timeout 300 python scripts/bench_lstm.py > $OUT/bench_lstm_r02c.log 2>&1; echo "bench_lstm rc=$?"
timeout 200 python scripts/check_lstm_tc.py trace bwd > $OUT/trace_r02c.log 2>&1; echo "trace rc=$?"
grep -h "lstm_seq\|median\|period" $OUT/bench_lstm_r02c.log $OUT/trace_r02c.log | cut -c1-400
