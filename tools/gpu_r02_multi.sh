#!/bin/bash
# Round-2 multi-GPU visit (gpurun --gpus N): the C-ABI peer path across processes (IPC) and inside one process
# (OptixHello --gpus), checked bit for bit, then bench.py at N (device and host consumers, secondary config 5).
set -u
N=${1:-2}
TAG=${2:-r02m$N}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $OUT/smi.txt 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $RUN tools/check_bands_gpu.py > $OUT/bands_check.log 2>&1; echo "bands check exit: $?" >> $OUT/bands_check.log
timeout 300 python -m pytest tests/test_peer_frames_gpu.py -m gpu -q -x > $OUT/pytest_peer.log 2>&1; echo "pytest exit: $?" >> $OUT/pytest_peer.log
for g in 1 $N; do
  timeout 120 raytracingdiffusioncurves_b200/OptixHello tests/golden/xmls/arch.xml 128 --width 1920 --height 1080 --frames 50 --gpus $g --units-per-tile 4 --dump-f32 $OUT/arch_$g.f32 > $OUT/optixhello_arch_$g.log 2>&1
  timeout 120 raytracingdiffusioncurves_b200/OptixHello tests/golden/xmls/DiffusionCurvePack/lady_bug.xml 128 --width 1920 --height 1080 --frames 20 --gpus $g --units-per-tile 1 --dump-f32 $OUT/lady_$g.f32 > $OUT/optixhello_lady_$g.log 2>&1
done
cmp $OUT/arch_1.f32 $OUT/arch_$N.f32 && echo "OptixHello arch: 1 GPU == $N GPUs" > $OUT/optixhello_cmp.log
cmp $OUT/lady_1.f32 $OUT/lady_$N.f32 && echo "OptixHello lady_bug: 1 GPU == $N GPUs" >> $OUT/optixhello_cmp.log
rm -f $OUT/*.f32
NCCL_DEBUG=INFO timeout 900 $RUN bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench exit: $?" >> $OUT/bench_n$N.err
timeout 600 $RUN bench.py --gpus $N --steps 10 --warmup 3 --workload ladybug_1080p_128rpp > $OUT/bench_ladybug_n$N.json 2>> $OUT/bench_n$N.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_n1.json 2> $OUT/bench_n1.err
cat $OUT/bands_check.log | tail -20; tail -3 $OUT/pytest_peer.log; cat $OUT/optixhello_cmp.log; grep -h "Average frame" $OUT/optixhello_*.log; tail -3 $OUT/bench_n$N.err | cut -c1-300
for f in $OUT/bench_n1.json $OUT/bench_n$N.json $OUT/bench_ladybug_n$N.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); s=d.get('secondary') or {}
    print(sys.argv[1].split('/')[-1], 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],3), 'kernel_ms', round((d.get('roofline') or {}).get('kernel_ms',0),3), 'secondary', round(s.get('value',0),2), round(s.get('ms_per_step',0),1), d['e2e'].get('host_frame_complete'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
