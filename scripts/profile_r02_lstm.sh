#!/bin/bash
# app-range replay of ONE persistent recurrence launch (the whole application is re-run per pass; only the range is measured)
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${TAG:-r02}
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__cycles_active.avg,l1tex__m_xbar2l1tex_read_bytes.sum,smsp__inst_executed.sum
for H in ${AVC_NCU_HS:-1024 512}; do
for which in fwd bwd; do
  AVC_NCU_H=$H AVC_NCU_WHICH=$which timeout 400 ncu --replay-mode app-range --clock-control none --metrics $M --csv --log-file $OUT/ncu_${TAG}_lstm_apprange_${which}_h$H.csv python scripts/ncu_lstm_range.py > $OUT/ncu_${TAG}_lstm_apprange_${which}_h$H.log 2>&1
  echo "app-range $which H=$H rc=$?"; grep -c . $OUT/ncu_${TAG}_lstm_apprange_${which}_h$H.csv
done
done
