#!/bin/bash
# Round-2 GPU visit A: all parity tests (incl. the full-size ones), smoke, bench N=1, then ncu captures of the SHIPPED
# k_render on the headline frame, a band of config 5 and a tree scene (each after its plain run exited 0).
set -u
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
nproc > $OUT/nproc.txt
python -m pytest tests -m gpu -q -s --durations=15 > $OUT/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> $OUT/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit: $?" >> $OUT/smoke.log
python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit: $?" >> $OUT/bench.err
prof() {  # name, env assignments..., then "--", then the command
  local name=$1; shift
  env "$@" > $OUT/plain_$name.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render --launch-skip 1 -c 1 -f -o $OUT/prof_$name \
    env "$@" > $OUT/ncu_$name.log 2>&1
  echo "$name: ncu exit $?" >> $OUT/ncu_status.log
}
prof arch RDC_X=1 python tools/profile_frame.py arch_1080p_128rpp 2
prof synth8k_band RDC_PROFILE_ROWS=4096:4352 python tools/profile_frame.py synth100k_8k_512rpp 2
prof ladybug RDC_PROFILE_SIZE=1920x1080x64 python tools/profile_frame.py ladybug_1080p_128rpp 2
prof dolphin RDC_PROFILE_SIZE=1920x1080x64 python tools/profile_frame.py dolphin_4k_256rpp 2
python tools/profile_frame.py arch_1080p_128rpp 3 > $OUT/profile_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv \
    python tools/profile_frame.py arch_1080p_128rpp 3 > $OUT/ncu_launches.log 2>&1
python tools/sweep_scenes.py 3840 2160 256 --modes > $OUT/sweep_4k_256rpp.jsonl 2>> $OUT/bench.err
for wl in portal_1080p_depth31 ladybug_1080p_128rpp dolphin_4k_256rpp synth100k_2k_64rpp; do
  python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline > $OUT/bench_$wl.json 2>> $OUT/bench.err
done
python bench.py --steps 2 --warmup 3 --workload synth100k_8k_512rpp --no-cpu-baseline > $OUT/bench_synth100k_8k_512rpp.json 2>> $OUT/bench.err
tail -5 $OUT/pytest_gpu.log; cat $OUT/smoke.log; cat $OUT/ncu_status.log; tail -3 $OUT/bench.err; tail -1 $OUT/sweep_4k_256rpp.jsonl
