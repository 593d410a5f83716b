#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for tree in auto radix median; do
for wl in arch_1080p_128rpp ladybug_1080p_128rpp portal_1080p_depth31 dolphin_4k_256rpp synth100k_2k_64rpp; do echo "== $tree $wl"; if [ $tree = auto ]; then python tools/profile_frame.py $wl 3 2>&1 | tail -1; else RDC_B200_TREE=$tree python tools/profile_frame.py $wl 3 2>&1 | tail -1; fi; done; done | tee gpurun_out/tree.log
