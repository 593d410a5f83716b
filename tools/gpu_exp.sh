#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for rl in 0 1 2 4 8; do
  for wl in arch_1080p_128rpp ladybug_1080p_128rpp dolphin_4k_256rpp portal_1080p_depth31 synth100k_2k_64rpp; do
    echo "== run_length=$rl $wl"
    RDC_RUN_LENGTH=$rl python tools/profile_frame.py $wl 3 2>&1 | tail -1
  done
done 2>&1 | tee gpurun_out/runlen.log
