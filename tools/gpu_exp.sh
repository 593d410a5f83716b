#!/bin/bash
# kernel experiment: parity first, then frame times of library variants on several workloads
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for lib in "" build/librdc_b200_mb3.so build/librdc_b200_mb2.so; do
  for wl in arch_1080p_128rpp ladybug_1080p_128rpp portal_1080p_depth31 synth100k_2k_64rpp; do
    echo "== lib=${lib:-default} $wl"
    RDC_B200_LIB=${lib:+$PWD/$lib} python tools/profile_frame.py $wl 4 2>&1 | tail -2
  done
done 2>&1 | tee gpurun_out/variants.log
