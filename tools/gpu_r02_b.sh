#!/bin/bash
# Round-2 GPU visit B (one GPU): all parity tests incl. the peer-frame tests, bench with the secondary block, kernel
# variants (slab in FMA form, while-while traversal) A/B on headline / tree scenes / config-5 band, screencap renders.
set -u
TAG=${1:-r02b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $OUT/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit: $?" >> $OUT/bench.err
for v in shipped slab ww slabww; do
  lib=build/librdc_b200_$v.so; [ $v = shipped ] && lib=raytracingdiffusioncurves_b200/librdc_b200.so
  {
    echo "== $v"
    RDC_B200_LIB=$lib python -m pytest tests -m gpu -x -q -k "golden or small or switches or portals or local or synthetic or config3 or config4 or config5 or headline" 2>&1 | tail -2
    for wl in arch_1080p_128rpp portal_1080p_depth31 ladybug_1080p_128rpp; do
      echo "-- $wl: $(RDC_B200_LIB=$lib python tools/profile_frame.py $wl 4 2>&1 | tail -1)"
    done
    echo "-- dolphin 4k@256: $(RDC_B200_LIB=$lib python tools/profile_frame.py dolphin_4k_256rpp 3 2>&1 | tail -1)"
    echo "-- synth 8k@512 rows 4096:4352: $(RDC_B200_LIB=$lib RDC_PROFILE_ROWS=4096:4352 python tools/profile_frame.py synth100k_8k_512rpp 2 2>&1 | tail -1)"
    RDC_B200_LIB=$lib python tools/sweep_scenes.py 3840 2160 256 > $OUT/sweep_$v.jsonl 2>&1
    tail -1 $OUT/sweep_$v.jsonl
  } >> $OUT/variants.log 2>&1
done
echo "== headline through the local run table (route 2)" >> $OUT/variants.log
RDC_PROFILE_ROUTE=2 RDC_PROFILE_STATS=1 python tools/profile_frame.py arch_1080p_128rpp 4 >> $OUT/variants.log 2>&1
# renders for the comparison with the reference's screenshots (tools/rdc_diff.py runs where /root/reference is)
for s in DiffusionCurvePack/lady_bug endcap weight_demo; do
  b=$(basename $s)
  raytracingdiffusioncurves_b200/OptixHello tests/golden/xmls/$s.xml 128 --out $OUT/$b.png --dump-f32 $OUT/$b.f32 > $OUT/optixhello_$b.log 2>&1
done
raytracingdiffusioncurves_b200/OptixHello tests/golden/xmls/DiffusionCurvePack/lady_bug.xml 16 --frames 12 --accumulate --scroll-at 6:1 --out $OUT/lady_bug_acc.png > $OUT/optixhello_acc.log 2>&1
tail -4 $OUT/pytest_gpu.log; tail -2 $OUT/bench.err; cat $OUT/variants.log
