#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
for wl in arch_1080p_128rpp ladybug_1080p_128rpp dolphin_4k_256rpp; do echo "== $wl"; python tools/profile_frame.py $wl 3 2>&1 | tail -1; done | tee gpurun_out/head6.log
RDC_PROFILE_ROWS=4096:4352 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render --launch-skip 1 -c 1 \
    -f -o gpurun_out/prof_local_synth8k_band python tools/profile_frame.py synth100k_8k_512rpp 2 > gpurun_out/ncu_local_synth8k.log 2>&1
echo "ncu exit $?"
