L=gpurun_out/ws3.log
echo "== default" > $L; timeout 200 python scripts/check_lstm_ws.py big 2>&1 | grep -o '"B": [0-9]*, "T": [0-9]*, "H": [0-9]*\|"us_ring.*\|FAILS.*\|rror.*' >> $L
for k in 2 1; do echo "== KBP=$k" >> $L; AVC_LSTM_WS_KBP=$k timeout 120 python scripts/check_lstm_ws.py big 2>&1 | grep -o '"B": [0-9]*, "T": [0-9]*, "H": [0-9]*\|"us_ring.*\|FAILS.*\|rror.*' >> $L; done
echo "== tests" >> $L
timeout 900 python -m pytest tests/test_gpu_lstm_tc.py tests/test_gpu_determinism.py tests/test_gpu_benched_shapes.py tests/test_gpu_dvector.py tests/test_gpu_conversion.py -x -q 2>&1 | tail -15 >> $L
echo "== bench" >> $L
timeout 600 python bench.py > gpurun_out/ws3_bench.json 2>> $L; echo rc=$? >> $L
cat $L; python -c "
import json; d=json.loads(open('gpurun_out/ws3_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['families'].items()})"
