#!/bin/bash
# ncu --set full of k_render on the tree-path workloads (one launch each). Reports land in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-tree}
for spec in synth100k_2k_64rpp:2048x2048x64 dolphin_4k_256rpp:1920x1080x64 ladybug_1080p_128rpp:1920x1080x64; do
  wl=${spec%%:*}; size=${spec##*:}
  RDC_PROFILE_SIZE=$size python tools/profile_frame.py $wl 3 > gpurun_out/plain_${TAG}_$wl.log 2>&1 || { echo "plain run failed $wl"; continue; }
  tail -1 gpurun_out/plain_${TAG}_$wl.log
  RDC_PROFILE_SIZE=$size timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render --launch-skip 1 -c 1 \
    -f -o gpurun_out/prof_${TAG}_$wl python tools/profile_frame.py $wl 2 > gpurun_out/ncu_${TAG}_$wl.log 2>&1
  echo "ncu exit $?"
done
