#!/bin/bash
# Round-2 GPU visit K: units per tile for cut-table launches at 1080p (whole frame and a rank's share).
set -u
OUT=gpurun_out/${1:-r02k}; mkdir -p $OUT
{
for wl in ladybug_1080p_128rpp; do
  for share in 0:0 8:0; do
    for units in 1 2 4 8; do
      echo "== $wl strips $share units $units: $(RDC_PROFILE_STRIPS=$share RDC_PROFILE_UNITS=$units python tools/profile_frame.py $wl 6 2>&1 | tail -1)"
    done
  done
done
for units in 1 2 4; do echo "== dolphin 1080p@128 units $units: $(RDC_PROFILE_SIZE=1920x1080x128 RDC_PROFILE_UNITS=$units python tools/profile_frame.py dolphin_4k_256rpp 5 2>&1 | tail -1)"; done
for units in 1 2 4; do echo "== face-like: behindthecurtain n/a"; done | head -0
for i in 1 2 3; do echo "== arch 1080p run $i: $(python tools/profile_frame.py arch_1080p_128rpp 8 2>&1 | tail -1)"; done
} > $OUT/units.log 2>&1
cat $OUT/units.log
