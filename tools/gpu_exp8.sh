#!/bin/bash
set -u
mkdir -p gpurun_out
{
for thr in 65 1024; do echo "== local_min_runs $thr, 4k@256"; RDC_B200_LOCAL_MIN_RUNS=$thr python tools/sweep_scenes.py 3840 2160 256 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    if d.get('runs',0) > 64 or 'total_ms' in d: print(d.get('scene','TOTAL'), d.get('runs'), d.get('render_ms', d.get('total_ms')))
"; done
echo "== local_min_runs 65, 1080p@128"; RDC_B200_LOCAL_MIN_RUNS=65 python tools/sweep_scenes.py 1920 1080 128 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    if d.get('runs',0) > 64 or 'total_ms' in d: print(d.get('scene','TOTAL'), d.get('runs'), d.get('render_ms', d.get('total_ms')))
"
echo "== default, 1080p@128"; python tools/sweep_scenes.py 1920 1080 128 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    if d.get('runs',0) > 64 or 'total_ms' in d: print(d.get('scene','TOTAL'), d.get('runs'), d.get('render_ms', d.get('total_ms')))
"
} 2>&1 | tee gpurun_out/exp8.log
