#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
export RDC_B200_LOCAL_MIN_RUNS=65
python tools/sweep_scenes.py 1920 1080 64 --modes 2>&1 | grep -E "Pack|test4|test5|frame" | tee gpurun_out/sweep_1080p_64_modes.log
RDC_RUN_LENGTH=8 python tools/sweep_scenes.py 1920 1080 64 --modes 2>&1 | grep -E "Pack|frame" | tee gpurun_out/sweep_1080p_64_modes_rl8.log
export RDC_PROFILE_STATS=1
{
for e in "A=1" "RDC_RUN_LENGTH=8" "RDC_RUN_LENGTH=2"; do
echo "== synth 8k@512 rows 4096:4352 [$e]"; env $e RDC_PROFILE_ROWS=4096:4352 python tools/profile_frame.py synth100k_8k_512rpp 2 2>&1 | tail -2
echo "== synth 2k@64 [$e]"; env $e python tools/profile_frame.py synth100k_2k_64rpp 2 2>&1 | tail -2
done
} 2>&1 | tee gpurun_out/head4.log
