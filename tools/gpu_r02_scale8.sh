#!/bin/bash
# Round-2 scaling visit on an 8-GPU box: bands check + OptixHello at 8, then bench.py at N = 1, 4, 8 (headline + secondary).
set -u
TAG=${1:-r02s8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $OUT/smi.txt 2>&1
run() { local n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 "$@"; }
run 8 tools/check_bands_gpu.py > $OUT/bands_check_8.log 2>&1; echo "bands check exit: $?" >> $OUT/bands_check_8.log
for g in 1 8; do
  timeout 120 raytracingdiffusioncurves_b200/OptixHello tests/golden/xmls/arch.xml 128 --width 1920 --height 1080 --frames 50 --gpus $g --units-per-tile 4 --dump-f32 $OUT/arch_$g.f32 > $OUT/optixhello_arch_$g.log 2>&1
done
cmp $OUT/arch_1.f32 $OUT/arch_8.f32 && echo "OptixHello arch: 1 GPU == 8 GPUs" > $OUT/optixhello_cmp.log
rm -f $OUT/*.f32
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/scale_arch_n1.json 2> $OUT/scale_n1.err
for n in 4 8; do
  run $n bench.py --gpus $n --steps 20 --warmup 5 > $OUT/scale_arch_n$n.json 2> $OUT/scale_n$n.err; echo "bench N=$n exit: $?" >> $OUT/scale_n$n.err
done
run 8 bench.py --gpus 8 --steps 10 --warmup 3 --workload ladybug_1080p_128rpp > $OUT/scale_ladybug_n8.json 2>> $OUT/scale_n8.err
tail -6 $OUT/bands_check_8.log; cat $OUT/optixhello_cmp.log; grep -h "Average frame" $OUT/optixhello_*.log
for f in $OUT/scale_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); s=d.get('secondary') or {}
    print(sys.argv[1].split('/')[-1], 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],4), 'kernel_ms', round((d.get('roofline') or {}).get('kernel_ms',0),4), 'secondary', round(s.get('value',0),2), round(s.get('ms_per_step',0),1), d['e2e'].get('host_frame_complete'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
