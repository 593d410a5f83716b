#!/bin/bash
# Quick GPU visit: parity tests, then one-line timings of the bench workloads (tools/profile_frame.py).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
for wl in arch_1080p_128rpp arch_512_128rpp portal_1080p_depth31 ladybug_1080p_128rpp synth100k_2k_64rpp; do echo "== $wl"; python tools/profile_frame.py $wl 3 2>&1 | tail -1; done | tee gpurun_out/head.log
echo "== dolphin 1080p@64"; RDC_PROFILE_SIZE=1920x1080x64 python tools/profile_frame.py dolphin_4k_256rpp 3 2>&1 | tail -1 | tee -a gpurun_out/head.log
