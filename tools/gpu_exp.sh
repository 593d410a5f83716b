#!/bin/bash
set -u
mkdir -p gpurun_out
for sp in 4 2 1; do
  for wl in arch_1080p_128rpp ladybug_1080p_128rpp portal_1080p_depth31 synth100k_2k_64rpp; do
    echo "== split=$sp $wl"
    RDC_B200_SPLIT=$sp python tools/profile_frame.py $wl 3 2>&1 | tail -1
  done
done 2>&1 | tee gpurun_out/split.log
for wl in portal_1080p_depth31 ladybug_1080p_128rpp; do
  echo "== mb3 $wl"; RDC_B200_LIB=$PWD/build/librdc_b200_mb3.so python tools/profile_frame.py $wl 3 2>&1 | tail -1
done 2>&1 | tee -a gpurun_out/split.log
