#!/bin/bash
# Round-2 GPU visit J (one GPU): cut table refined per tile — parity, every route on every scene, config-5 band on both tables.
set -u
TAG=${1:-r02j}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest exit: $?" >> $OUT/pytest_gpu.log
python tools/sweep_scenes.py 3840 2160 256 --modes > $OUT/sweep_modes.jsonl 2>&1
{
for wl in ladybug_1080p_128rpp portal_1080p_depth31 arch_1080p_128rpp; do echo "-- $wl: $(RDC_PROFILE_STATS=1 python tools/profile_frame.py $wl 4 2>&1 | tail -3 | head -2 | tr '\n' ' ')"; done
echo "-- dolphin 4k: $(RDC_PROFILE_STATS=1 python tools/profile_frame.py dolphin_4k_256rpp 3 2>&1 | tail -3 | head -2 | tr '\n' ' ')"
echo "-- ladybug strips 8:0: $(RDC_PROFILE_STRIPS=8:0 python tools/profile_frame.py ladybug_1080p_128rpp 4 2>&1 | tail -1)"
echo "-- synth band, local table (Morton tree): $(RDC_PROFILE_STATS=1 RDC_PROFILE_ROWS=4096:4352 python tools/profile_frame.py synth100k_8k_512rpp 2 2>&1 | tail -3 | head -2 | tr '\n' ' ')"
echo "-- synth band, cut table (SAH tree forced): $(RDC_TREE=2 RDC_PROFILE_STATS=1 RDC_PROFILE_ROWS=4096:4352 python tools/profile_frame.py synth100k_8k_512rpp 2 2>&1 | tail -3 | head -2 | tr '\n' ' ')"
echo "-- synth 2k@64, local: $(python tools/profile_frame.py synth100k_2k_64rpp 3 2>&1 | tail -1)"
echo "-- synth 2k@64, cut (SAH forced): $(RDC_TREE=2 python tools/profile_frame.py synth100k_2k_64rpp 3 2>&1 | tail -1)"
} > $OUT/cut.log 2>&1
tail -3 $OUT/pytest_gpu.log; cat $OUT/cut.log; tail -1 $OUT/sweep_modes.jsonl
python - $OUT/sweep_modes.jsonl <<'PY'
import json,sys
a=[json.loads(l) for l in open(sys.argv[1]) if l.startswith('{"scene')]
for x in a: print(f"{x['scene']:45s} runs {x['runs']:5d} auto {x['render_ms']:8.2f} tree {x['tree_render_ms']:8.2f} local {x['local_render_ms']:8.2f} cut {x['cut_render_ms']:8.2f}")
PY
