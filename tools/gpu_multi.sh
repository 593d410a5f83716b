#!/bin/bash
# multi-GPU visit: band equivalence check, then bench at N=1 and N=G
set -u
G=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511"
$TR tools/check_bands_gpu.py 2>&1 | grep -E "OK|FAIL|Error|error" | tee gpurun_out/bands_check_$G.log
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err
$TR bench.py --gpus $G --steps 100 --warmup 5 > gpurun_out/bench_n$G.json 2> gpurun_out/bench_n$G.err; tail -3 gpurun_out/bench_n$G.err
$TR bench.py --gpus $G --steps 10 --warmup 3 --workload ladybug_1080p_128rpp > gpurun_out/bench_ladybug_n$G.json 2>> gpurun_out/bench_n$G.err
$TR bench.py --gpus $G --steps 5 --warmup 3 --workload synth100k_2k_64rpp > gpurun_out/bench_synth2k_n$G.json 2>> gpurun_out/bench_n$G.err
python bench.py --steps 10 --warmup 3 --workload ladybug_1080p_128rpp --no-cpu-baseline > gpurun_out/bench_ladybug_n1.json 2>> gpurun_out/bench_n1.err
python bench.py --steps 5 --warmup 3 --workload synth100k_2k_64rpp --no-cpu-baseline > gpurun_out/bench_synth2k_n1.json 2>> gpurun_out/bench_n1.err
for f in gpurun_out/bench_n1.json gpurun_out/bench_n$G.json gpurun_out/bench_ladybug_n1.json gpurun_out/bench_ladybug_n$G.json gpurun_out/bench_synth2k_n1.json gpurun_out/bench_synth2k_n$G.json; do
python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], 'n', d['n_gpus'], round(d['value'],2), 'Grays/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['ms_per_step'],3), 'ms', 'frac', d['roofline'] and round(d['roofline']['frac'],3), d['clocks'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
