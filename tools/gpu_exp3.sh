#!/bin/bash
set -u
mkdir -p gpurun_out
export RDC_PROFILE_STATS=1
{
for e in "A=1" "RDC_RUN_LENGTH=8" "RDC_RUN_LENGTH=4" "RDC_B200_NO_LOCAL=1"; do
echo "== synth 8k@512 rows 4096:4352 [$e]"; env $e RDC_PROFILE_ROWS=4096:4352 python tools/profile_frame.py synth100k_8k_512rpp 2 2>&1 | tail -2
done
for e in "A=1" "RDC_RUN_LENGTH=8" "RDC_B200_NO_LOCAL=1"; do
echo "== dolphin 4k@256 [$e]"; env $e python tools/profile_frame.py dolphin_4k_256rpp 2 2>&1 | tail -2
done
} 2>&1 | tee gpurun_out/head3.log
unset RDC_PROFILE_STATS
RDC_RUN_LENGTH=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render --launch-skip 1 -c 1 \
    -f -o gpurun_out/prof_local_synth2k python tools/profile_frame.py synth100k_2k_64rpp 2 > gpurun_out/ncu_local_synth2k.log 2>&1
echo "ncu exit $?"
