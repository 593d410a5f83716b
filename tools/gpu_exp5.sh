#!/bin/bash
set -u
mkdir -p gpurun_out
{
for sp in 4 2 1; do echo "== split $sp, 4k@256"; RDC_B200_SPLIT=$sp python tools/sweep_scenes.py 3840 2160 256 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d.get('scene','TOTAL'), d.get('render_ms', d.get('total_ms')))
"; done
for lim in 40960 57344 73728; do echo "== smem limit $lim, 1080p@128"; RDC_B200_SMEM_LIMIT=$lim python tools/sweep_scenes.py 1920 1080 128 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    if 'Pack' in d.get('scene','') or 'total_ms' in d or 'test4' in d.get('scene','') or 'test5' in d.get('scene',''): print(d.get('scene','TOTAL'), d.get('render_ms', d.get('total_ms')))
"; done
for sp in 4 2 1; do echo "== split $sp, synth 8k@512 band"; RDC_B200_SPLIT=$sp RDC_PROFILE_ROWS=4096:4352 python tools/profile_frame.py synth100k_8k_512rpp 2 2>&1 | tail -1; done
} 2>&1 | tee gpurun_out/exp5.log
