#!/bin/bash
# Round-2 GPU visit C (one GPU): where does a rank's launch lose time at 8 GPUs? One rank's share of the headline frame
# (strips 8:0) with the launch timeline of the counting build, for 1/2/4 units per tile; then the parity suite.
set -u
TAG=${1:-r02c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
{
for share in 0:0 2:0 4:0 8:0 8:3; do
  for units in 0 2 1; do
    echo "== strips $share units/tile $units"
    RDC_PROFILE_STRIPS=$share RDC_PROFILE_UNITS=$units RDC_PROFILE_STATS=1 python tools/profile_frame.py arch_1080p_128rpp 6 2>&1 | tail -4
  done
done
echo "== lady_bug strips 8:0"; RDC_PROFILE_STRIPS=8:0 RDC_PROFILE_STATS=1 python tools/profile_frame.py ladybug_1080p_128rpp 4 2>&1 | tail -4
echo "== lady_bug full"; RDC_PROFILE_STATS=1 python tools/profile_frame.py ladybug_1080p_128rpp 4 2>&1 | tail -4
} > $OUT/timeline.log 2>&1
python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest exit: $?" >> $OUT/pytest_gpu.log
cat $OUT/timeline.log; tail -3 $OUT/pytest_gpu.log
