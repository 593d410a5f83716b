#!/bin/bash
# Round-2 GPU visit G (one GPU): surface-area-heuristic tree (host build) against the Morton radix tree.
set -u
TAG=${1:-r02g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -m gpu -q -x -s -k "tree_builders" > $OUT/pytest_trees.log 2>&1; echo "pytest exit: $?" >> $OUT/pytest_trees.log
python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest exit: $?" >> $OUT/pytest_gpu.log
{
for tree in 1 2; do
  echo "== tree $tree (1 Morton, 2 SAH)"
  RDC_TREE=$tree python tools/sweep_scenes.py 3840 2160 256 > $OUT/sweep_tree$tree.jsonl 2>&1; tail -1 $OUT/sweep_tree$tree.jsonl
  for wl in ladybug_1080p_128rpp portal_1080p_depth31; do echo "-- $wl: $(RDC_TREE=$tree RDC_PROFILE_STATS=1 python tools/profile_frame.py $wl 4 2>&1 | tail -3 | head -2 | tr '\n' ' ')"; done
  echo "-- dolphin 4k: $(RDC_TREE=$tree RDC_PROFILE_STATS=1 python tools/profile_frame.py dolphin_4k_256rpp 3 2>&1 | tail -3 | head -2 | tr '\n' ' ')"
done
echo "== synth (auto = Morton above 65536 runs) band: $(RDC_PROFILE_ROWS=4096:4352 python tools/profile_frame.py synth100k_8k_512rpp 2 2>&1 | tail -1)"
echo "== synth SAH forced band: $(RDC_TREE=2 RDC_PROFILE_ROWS=4096:4352 python tools/profile_frame.py synth100k_8k_512rpp 2 2>&1 | tail -1)"
} > $OUT/trees.log 2>&1
grep -h "boxes tested\|passed\|failed" $OUT/pytest_trees.log; tail -2 $OUT/pytest_gpu.log; cat $OUT/trees.log
python - $OUT/sweep_tree1.jsonl $OUT/sweep_tree2.jsonl <<'PY'
import json,sys
a=[json.loads(l) for l in open(sys.argv[1]) if l.startswith('{"scene')]; b=[json.loads(l) for l in open(sys.argv[2]) if l.startswith('{"scene')]
for x,y in zip(a,b): print(f"{x['scene']:45s} runs {x['runs']:5d} morton {x['render_ms']:8.2f} sah {y['render_ms']:8.2f}")
PY
