#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for wl in arch_1080p_128rpp ladybug_1080p_128rpp portal_1080p_depth31 synth100k_2k_64rpp; do
  python bench.py --steps 5 --warmup 3 --workload $wl --no-cpu-baseline 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$wl', round(d['ms_per_step'],3),'ms', d['config']['chords'],'chords depth',d['config']['bvh_depth'], d['roofline']['per_ray'], 'kernel_ms', round(d['roofline']['kernel_ms'],3),'e2e ms', d['e2e']['ms_per_step'])
    else: print(l.rstrip())
"
done 2>&1 | tee gpurun_out/stats.log
