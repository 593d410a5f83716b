#!/bin/bash
# A/B visit for kernel variants built here with `make variantd NAME=<n> DEFS=...` (build/librdc_b200_<n>.so travels with
# the snapshot): for the shipped library and every variant named on the command line, a parity subset and the frame
# times of five workloads. Output: gpurun_out/variants/<name>.log
#   tools/gpu_variants.sh recp w4 norec
set -u
OUT=gpurun_out/variants
mkdir -p $OUT
run_one() {  # name, library path
  local name=$1 lib=$2
  {
    echo "=== $name ($lib)"
    RDC_B200_LIB=$lib python -m pytest tests -m gpu -x -q -k "golden or small or switches or portals or local or synthetic" 2>&1 | tail -2
    for wl in arch_1080p_128rpp portal_1080p_depth31 ladybug_1080p_128rpp synth100k_2k_64rpp; do
      echo "-- $wl: $(RDC_B200_LIB=$lib python tools/profile_frame.py $wl 4 2>&1 | tail -1)"
    done
    echo "-- dolphin 4k@256: $(RDC_B200_LIB=$lib python tools/profile_frame.py dolphin_4k_256rpp 3 2>&1 | tail -1)"
    echo "-- synth 8k@512 rows 4096:4352: $(RDC_B200_LIB=$lib RDC_PROFILE_ROWS=4096:4352 python tools/profile_frame.py synth100k_8k_512rpp 2 2>&1 | tail -1)"
  } 2>&1 | tee $OUT/$name.log
}
run_one shipped raytracingdiffusioncurves_b200/librdc_b200.so
for v in "$@"; do
  if [ -f build/librdc_b200_$v.so ]; then run_one $v build/librdc_b200_$v.so; else echo "no build/librdc_b200_$v.so (make variantd NAME=$v DEFS=...)"; fi
done
