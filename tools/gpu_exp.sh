#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for wl in arch_1080p_128rpp arch_512_128rpp portal_1080p_depth31 ladybug_1080p_128rpp; do echo "== table $wl"; python tools/profile_frame.py $wl 3 2>&1 | tail -1; echo "== tree $wl"; RDC_B200_NO_TABLE=1 python tools/profile_frame.py $wl 3 2>&1 | tail -1; done | tee gpurun_out/table.log
python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['per_ray'], d['roofline']['frac'])"
