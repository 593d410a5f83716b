#!/bin/bash
# 8-GPU visit: strip equivalence at 8 ranks, headline bench at N=1,2,4,8, synthetic scene scaling
set -u
mkdir -p gpurun_out
run() { n=$1; shift; if [ "$n" = 1 ]; then python bench.py --gpus 1 "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@"; fi; }
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/check_bands_gpu.py 2>&1 | grep -E "OK|FAIL|rror" | tee gpurun_out/bands_check_8.log
for n in 1 2 4 8; do run $n --steps 100 --warmup 5 --no-cpu-baseline 2>gpurun_out/scale_n$n.err | grep '^{' > gpurun_out/scale_arch_n$n.json; done
for n in 1 8; do run $n --steps 5 --warmup 3 --no-cpu-baseline --workload synth100k_2k_64rpp 2>>gpurun_out/scale_n$n.err | grep '^{' > gpurun_out/scale_synth2k_n$n.json; done
for n in 1 8; do run $n --steps 5 --warmup 3 --no-cpu-baseline --workload ladybug_1080p_128rpp 2>>gpurun_out/scale_n$n.err | grep '^{' > gpurun_out/scale_ladybug_n$n.json; done
run 8 --steps 2 --warmup 1 --no-cpu-baseline --workload synth100k_8k_512rpp 2>>gpurun_out/scale_n8.err | grep '^{' > gpurun_out/scale_synth8k_n8.json
for f in gpurun_out/scale_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read()); print(sys.argv[1].split('/')[-1], 'n', d['n_gpus'], round(d['value'],2), 'Grays/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['ms_per_step'],3), 'ms', 'kernel', round(d['roofline']['kernel_ms'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
tail -3 gpurun_out/scale_n8.err
