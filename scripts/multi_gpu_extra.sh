set -u
OUT=gpurun_out; N=8; PORT=29700
run() { name=$1; shift; PORT=$((PORT + 1))
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N "$@" \
    > $OUT/bench_r02_n${N}_$name.json 2> $OUT/bench_r02_n${N}_$name.err || { echo "$name FAILED"; tail -5 $OUT/bench_r02_n${N}_$name.err; }
  python -c "
import json
d = json.loads(open('$OUT/bench_r02_n${N}_$name.json').read().strip().splitlines()[-1])
print('$name', 'n_gpus', d['n_gpus'], round(d['value'], 1), d['unit'], round(d['ms_per_step'], 3), 'ms/step')"
}
run c2_fp32 --precision fp32 --steps 10 --warmup 3 --sustained 0 --no-modes
run c5_convert --workload convert --utterances 4096 --steps 2 --warmup 3
