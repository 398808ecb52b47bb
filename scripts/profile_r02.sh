#!/bin/bash
# Round-2 profiler evidence (under gpurun, ONE GPU): launch list of a bench step, --set full of the 3xTF32 kernels, and a
# range-replay attempt on the persistent recurrences.
set -u
OUT=gpurun_out; mkdir -p $OUT
bash scripts/gpu_profile.sh r02 2>&1 | tail -3
python scripts/ncu_x3.py > /dev/null 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:tc_gemm_kernel|split3|lstm_gate|wgrad_reduce' -c 12 -f -o $OUT/prof_r02_x3 python scripts/ncu_x3.py > $OUT/ncu_r02_x3.log 2>&1
echo "x3 full rc=$?"
ncu -i $OUT/prof_r02_x3.ncu-rep --page raw --csv > $OUT/prof_r02_x3_raw.csv 2>/dev/null
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__cycles_active.avg
for which in fwd bwd; do
  AVC_NCU_WHICH=$which timeout 300 ncu --replay-mode range --clock-control none --metrics $M --csv --log-file $OUT/ncu_r02_lstm_range_$which.csv python scripts/ncu_lstm_range.py > $OUT/ncu_r02_lstm_range_$which.log 2>&1
  echo "range $which rc=$?"; tail -3 $OUT/ncu_r02_lstm_range_$which.log
done
ls -la $OUT/prof_r02_x3.ncu-rep $OUT/launches_r02.csv 2>&1 | cut -c20-120
