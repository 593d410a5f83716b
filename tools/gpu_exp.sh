#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for wl in arch_1080p_128rpp ladybug_1080p_128rpp portal_1080p_depth31 synth100k_2k_64rpp; do echo "== HEAD $wl"; python tools/profile_frame.py $wl 3 2>&1 | tail -1; done | tee gpurun_out/head.log
