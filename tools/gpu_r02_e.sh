#!/bin/bash
# Round-2 GPU visit E (one GPU): whole-scene table with distance-sorted slots — parity suite, table-mode workloads, bench.
set -u
TAG=${1:-r02e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest exit: $?" >> $OUT/pytest_gpu.log
{
for share in 0:0 2:0 4:0 8:0; do
  echo "== arch strips $share"; RDC_PROFILE_STRIPS=$share RDC_PROFILE_STATS=1 python tools/profile_frame.py arch_1080p_128rpp 6 2>&1 | tail -3
done
for wl in portal_1080p_depth31 ladybug_1080p_128rpp arch_512_128rpp; do echo "== $wl"; RDC_PROFILE_STATS=1 python tools/profile_frame.py $wl 4 2>&1 | tail -3; done
echo "== ladybug strips 8:0"; RDC_PROFILE_STRIPS=8:0 python tools/profile_frame.py ladybug_1080p_128rpp 4 2>&1 | tail -1
python tools/sweep_scenes.py 3840 2160 256 > $OUT/sweep_4k_256rpp.jsonl 2>&1; tail -1 $OUT/sweep_4k_256rpp.jsonl
} > $OUT/timeline.log 2>&1
python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit: $?" >> $OUT/bench.err
tail -3 $OUT/pytest_gpu.log; cat $OUT/timeline.log; tail -2 $OUT/bench.err
python - $OUT/bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); s=d.get('secondary') or {}
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],3), 'secondary', round(s.get('value',0),2))
PY
