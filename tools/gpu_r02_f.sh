#!/bin/bash
# Round-2 GPU visit F (one GPU): two variants that were written but never measured — 128-slot local run tables on the
# mid-size scenes (RDC_LOCAL_WORDS=4) and shading records for terminal hits of portal scenes.
set -u
TAG=${1:-r02f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
{
echo "== shipped, sweep with routes"; python tools/sweep_scenes.py 3840 2160 256 --modes > $OUT/sweep_shipped_modes.jsonl 2>&1; tail -1 $OUT/sweep_shipped_modes.jsonl
echo "== w4 (128 slots), sweep with routes"; RDC_B200_LIB=build/librdc_b200_w4.so python tools/sweep_scenes.py 3840 2160 256 --modes > $OUT/sweep_w4_modes.jsonl 2>&1; tail -1 $OUT/sweep_w4_modes.jsonl
echo "== w4 parity"; RDC_B200_LIB=build/librdc_b200_w4.so python -m pytest tests -m gpu -x -q -k "local or synthetic or config5 or config3" 2>&1 | tail -2
echo "== w4 synth band"; RDC_B200_LIB=build/librdc_b200_w4.so RDC_PROFILE_ROWS=4096:4352 python tools/profile_frame.py synth100k_8k_512rpp 2 2>&1 | tail -1
echo "== recp parity"; RDC_B200_LIB=build/librdc_b200_recp.so python -m pytest tests -m gpu -x -q -k "portal or config4 or golden" 2>&1 | tail -2
echo "== recp portal"; RDC_B200_LIB=build/librdc_b200_recp.so python tools/profile_frame.py portal_1080p_depth31 4 2>&1 | tail -1
echo "== shipped portal"; python tools/profile_frame.py portal_1080p_depth31 4 2>&1 | tail -1
} > $OUT/variants.log 2>&1
cat $OUT/variants.log
python - $OUT/sweep_shipped_modes.jsonl $OUT/sweep_w4_modes.jsonl <<'PY'
import json,sys
a=[json.loads(l) for l in open(sys.argv[1]) if l.startswith('{"scene')]; b=[json.loads(l) for l in open(sys.argv[2]) if l.startswith('{"scene')]
for x,y in zip(a,b): print(f"{x['scene']:45s} runs {x['runs']:5d} auto {x['render_ms']:8.2f} tree {x['tree_render_ms']:8.2f} local64 {x['local_render_ms']:8.2f} local128 {y['local_render_ms']:8.2f}")
PY
