#!/bin/bash
# Round-2 GPU visit O (one GPU): group mode for 2, 4 and 8 units per tile — parity, then the unit matrix again.
set -u
OUT=gpurun_out/${1:-r02o}; mkdir -p $OUT
python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest exit: $?" >> $OUT/pytest_gpu.log
{
for wl in arch_1080p_128rpp ladybug_1080p_128rpp; do
  for share in 0:0 2:0 4:0 8:0; do
    for units in 0 1 2 4 8; do
      echo "== $wl strips $share units $units: $(RDC_PROFILE_STRIPS=$share RDC_PROFILE_UNITS=$units python tools/profile_frame.py $wl 6 2>&1 | tail -1)"
    done
  done
done
echo "== arch 512 auto: $(python tools/profile_frame.py arch_512_128rpp 6 2>&1 | tail -1)"
} > $OUT/group.log 2>&1
tail -3 $OUT/pytest_gpu.log; cat $OUT/group.log
