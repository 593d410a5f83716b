#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
export RDC_PROFILE_STATS=1
run() { # label, env...
  echo "== $1"; shift
  env "$@" python tools/profile_frame.py ladybug_1080p_128rpp 3 2>&1 | tail -2
  env "$@" python tools/profile_frame.py synth100k_2k_64rpp 3 2>&1 | tail -2
  env "$@" RDC_PROFILE_SIZE=1920x1080x64 python tools/profile_frame.py dolphin_4k_256rpp 3 2>&1 | tail -2
}
{
run "default (64 slots)" A=1
run "64 slots, run length 8" RDC_RUN_LENGTH=8
run "128 slots" RDC_B200_LIB=build/librdc_b200_w4.so
run "128 slots, run length 8" RDC_B200_LIB=build/librdc_b200_w4.so RDC_RUN_LENGTH=8
run "tree" RDC_B200_NO_LOCAL=1
echo "== arch"; python tools/profile_frame.py arch_1080p_128rpp 3 2>&1 | tail -2
} 2>&1 | tee gpurun_out/head.log
