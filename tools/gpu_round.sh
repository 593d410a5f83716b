#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (all workloads), launch list, ncu capture, config-3 sweep, config 5.
# Most important first (a visit may be cut short by the GPU budget). Logs land in gpurun_out/<tag>/.
#   tools/gpu_round.sh <tag> [reference]     "reference" also runs the CPU arm (bench.py --impl reference)
set -u
TAG=${1:-round}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > $OUT/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> $OUT/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit: $?" >> $OUT/smoke.log
python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit: $?" >> $OUT/bench.err
ncu --set full --clock-control none --import-source on -k regex:k_render --launch-skip 1 -c 1 -f -o $OUT/prof_k_render_arch \
    python tools/profile_frame.py arch_1080p_128rpp 2 > $OUT/ncu_full.log 2>&1
python tools/profile_frame.py arch_1080p_128rpp 3 > $OUT/profile_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv \
    python tools/profile_frame.py arch_1080p_128rpp 3 > $OUT/ncu_launches.log 2>&1
for wl in arch_512_128rpp portal_1080p_depth31 ladybug_1080p_128rpp dolphin_4k_256rpp synth100k_2k_64rpp; do
  python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline > $OUT/bench_$wl.json 2>> $OUT/bench.err
done
python tools/sweep_scenes.py 3840 2160 256 > $OUT/sweep_4k_256rpp.jsonl 2>> $OUT/bench.err
python bench.py --steps 2 --warmup 3 --workload synth100k_8k_512rpp --no-cpu-baseline > $OUT/bench_synth100k_8k_512rpp.json 2>> $OUT/bench.err
if [ "${2:-}" = reference ]; then
  python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err
fi
tail -3 $OUT/pytest_gpu.log; cat $OUT/smoke.log; tail -3 $OUT/bench.err; tail -1 $OUT/sweep_4k_256rpp.jsonl
for f in $OUT/bench.json $OUT/bench_*rpp.json $OUT/bench_portal*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read()); r=d.get('roofline') or {}
    print(sys.argv[1].split('/')[-1], round(d['value'],3), 'Grays/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e'].get('ms_per_step',0),3), 'frac', round(r.get('frac',0),3), r.get('per_ray'), (d.get('cpu_baseline') or {}).get('value'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
