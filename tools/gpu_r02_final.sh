#!/bin/bash
# Round-2 final visit (one GPU): parity suite, smoke, bench (headline + secondary + CPU baseline), the other workloads,
# the 21-scene 4K sweep, then ncu captures of the SHIPPED k_render (each after its plain run exited 0) with the hash of the
# sources they were taken with, and the launch list.
set -u
TAG=${1:-r02final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
python -c "import bench; print(bench.source_sha16())" > $OUT/source_sha16.txt
python -m pytest tests -m gpu -q -s --durations=10 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit: $?" >> $OUT/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit: $?" >> $OUT/smoke.log
python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit: $?" >> $OUT/bench.err
for wl in arch_512_128rpp portal_1080p_depth31 ladybug_1080p_128rpp dolphin_4k_256rpp synth100k_2k_64rpp; do
  python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline > $OUT/bench_$wl.json 2>> $OUT/bench.err
done
python tools/sweep_scenes.py 3840 2160 256 > $OUT/sweep_4k_256rpp.jsonl 2>> $OUT/bench.err
prof() {
  local name=$1; shift
  env "$@" > $OUT/plain_$name.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render --launch-skip 1 -c 1 -f -o $OUT/prof_$name \
    env "$@" > $OUT/ncu_$name.log 2>&1
  echo "$name: ncu exit $?" >> $OUT/ncu_status.log
}
prof arch RDC_X=1 python tools/profile_frame.py arch_1080p_128rpp 2
prof synth8k_band RDC_PROFILE_ROWS=4096:4352 python tools/profile_frame.py synth100k_8k_512rpp 2
prof ladybug RDC_X=1 python tools/profile_frame.py ladybug_1080p_128rpp 2
prof dolphin RDC_PROFILE_SIZE=1920x1080x64 python tools/profile_frame.py dolphin_4k_256rpp 2
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/bench_plain_for_launches.json 2>> $OUT/bench.err &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/ncu_launches.log 2>&1
tail -4 $OUT/pytest_gpu.log; cat $OUT/smoke.log $OUT/ncu_status.log; tail -2 $OUT/bench.err; tail -1 $OUT/sweep_4k_256rpp.jsonl
for f in $OUT/bench.json $OUT/bench_*rpp.json $OUT/bench_portal*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d.get('roofline') or {}; s=d.get('secondary') or {}
    print(sys.argv[1].split('/')[-1], round(d['value'],2), 'Grays/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e'].get('ms_per_step',0),3), 'kernel', round(r.get('kernel_ms',0),3), 'frac', round(r.get('frac',0),3), 'l2', round((r.get('l2') or {}).get('frac',0),3), (d.get('cpu_baseline') or {}).get('value'), 'sec', round(s.get('value',0),2))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
