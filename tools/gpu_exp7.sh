#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
{
RDC_PROFILE_ROWS=4096:4352 python tools/profile_frame.py synth100k_8k_512rpp 2 2>&1 | tail -1
python tools/profile_frame.py synth100k_2k_64rpp 2 2>&1 | tail -1
python tools/profile_frame.py arch_1080p_128rpp 3 2>&1 | tail -1
python tools/sweep_scenes.py 3840 2160 256 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    if 'Pack' in d.get('scene','') or 'total_ms' in d: print(d.get('scene','TOTAL'), d.get('render_ms', d.get('total_ms')))
"
} 2>&1 | tee gpurun_out/exp7.log
