#!/bin/bash
# 8-GPU visit, short form: split equivalence at 8 ranks, then N=8 lines of the headline, lady_bug and config 5.
set -u
OUT=gpurun_out/scale8c
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$TR tools/check_bands_gpu.py 2>&1 | grep -E "OK|FAIL|SKIP|rror" | tee $OUT/bands_check_8.log
$TR bench.py --gpus 8 --steps 100 --warmup 5 --no-cpu-baseline 2>$OUT/scale_n8.err | grep '^{' > $OUT/scale_arch_n8.json
$TR bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --workload ladybug_1080p_128rpp 2>>$OUT/scale_n8.err | grep '^{' > $OUT/scale_ladybug_n8.json
$TR bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline --workload synth100k_8k_512rpp 2>>$OUT/scale_n8.err | grep '^{' > $OUT/scale_synth8k_n8.json
for f in $OUT/scale_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read()); print(sys.argv[1].split('/')[-1], 'n', d['n_gpus'], round(d['value'],2), 'Grays/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['ms_per_step'],3), 'ms', 'kernel', round(d['roofline']['kernel_ms'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
tail -3 $OUT/scale_n8.err
