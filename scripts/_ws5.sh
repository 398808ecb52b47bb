L=gpurun_out/ws5.log
echo "== fwd" > $L; timeout 200 python scripts/check_lstm_ws.py big 2>&1 | grep -o '"B": [0-9]*, "T": [0-9]*, "H": [0-9]*\|"us_ring.*\|FAILS.*\|rror.*' >> $L
echo "== bwd" >> $L; timeout 200 python scripts/check_lstm_ws.py bwd big 2>&1 | grep -o '"B": [0-9]*, "T": [0-9]*, "H": [0-9]*\|"us_ks.*\|FAILS.*\|rror.*' >> $L
echo "== tests" >> $L
timeout 2400 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_loss_curve.py 2>&1 | tail -8 >> $L
echo "== bench" >> $L
timeout 600 python bench.py > gpurun_out/ws5_bench.json 2>> $L; echo rc=$? >> $L
cat $L; python -c "
import json; d=json.loads(open('gpurun_out/ws5_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['families'].items()}, d['modes'])"
