#!/bin/bash
# A/B with repeats (process-to-process noise is several per cent): headline frame at 1 and 2 units per tile.
set -u
OUT=gpurun_out/${1:-r02p}; mkdir -p $OUT
{
for i in 1 2 3 4 5 6; do
  for units in 1 2; do
    echo "run $i units $units: $(RDC_PROFILE_UNITS=$units python tools/profile_frame.py arch_1080p_128rpp 8 2>&1 | tail -1)"
  done
done
for wl in ladybug_1080p_128rpp; do for share in 2:0 4:0 8:0; do echo "== $wl $share auto: $(RDC_PROFILE_STRIPS=$share python tools/profile_frame.py $wl 6 2>&1 | tail -1)"; done; done
} > $OUT/ab.log 2>&1
cat $OUT/ab.log
