#!/bin/bash
# multi-GPU visit: split-equivalence check (NCCL and peer-memory forms), then the headline bench at N=G both ways
set -u
G=${1:-2}
OUT=gpurun_out/multi$G
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511"
$TR tools/check_bands_gpu.py 2>&1 | grep -E "OK|FAIL|SKIP|Error|error" | tee $OUT/bands_check.log
$TR bench.py --gpus $G --steps 100 --warmup 5 --no-cpu-baseline > $OUT/bench_peer.json 2> $OUT/bench_peer.err; tail -3 $OUT/bench_peer.err
RDC_BENCH_NCCL=1 $TR bench.py --gpus $G --steps 100 --warmup 5 --no-cpu-baseline > $OUT/bench_nccl.json 2> $OUT/bench_nccl.err; tail -3 $OUT/bench_nccl.err
$TR bench.py --gpus $G --steps 10 --warmup 3 --workload ladybug_1080p_128rpp > $OUT/bench_ladybug_peer.json 2>> $OUT/bench_peer.err
RDC_BENCH_NCCL=1 $TR bench.py --gpus $G --steps 10 --warmup 3 --workload ladybug_1080p_128rpp > $OUT/bench_ladybug_nccl.json 2>> $OUT/bench_nccl.err
for f in $OUT/bench_*.json; do
python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], 'n', d['n_gpus'], round(d['value'],2), 'Grays/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['ms_per_step'],3), 'ms', d['config']['parallelism'][:110])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
