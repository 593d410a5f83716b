#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for wl in arch_1080p_128rpp arch_512_128rpp portal_1080p_depth31 ladybug_1080p_128rpp synth100k_2k_64rpp; do echo "== $wl"; python tools/profile_frame.py $wl 3 2>&1 | tail -1; done | tee gpurun_out/head.log
