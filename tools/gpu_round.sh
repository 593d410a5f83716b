#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, launch list. Logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?" >> gpurun_out/smoke.log
python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit: $?" >> gpurun_out/bench.err
python bench.py --steps 20 --warmup 3 --workload arch_512_128rpp --no-cpu-baseline > gpurun_out/bench_512.json 2>> gpurun_out/bench.err
python bench.py --steps 10 --warmup 3 --workload portal_1080p_depth31 --no-cpu-baseline > gpurun_out/bench_portal.json 2>> gpurun_out/bench.err
python bench.py --steps 5 --warmup 3 --workload synth100k_2k_64rpp --no-cpu-baseline > gpurun_out/bench_synth2k.json 2>> gpurun_out/bench.err
python tools/profile_frame.py arch_1080p_128rpp 3 > gpurun_out/profile_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv \
    python tools/profile_frame.py arch_1080p_128rpp 3 > gpurun_out/ncu_launches.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err; cat gpurun_out/profile_plain.log
