#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, launch list. Logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?" >> gpurun_out/smoke.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit: $?" >> gpurun_out/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
for wl in arch_512_128rpp portal_1080p_depth31 ladybug_1080p_128rpp dolphin_4k_256rpp synth100k_2k_64rpp; do
  python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline > gpurun_out/bench_$wl.json 2>> gpurun_out/bench.err
done
python tools/profile_frame.py arch_1080p_128rpp 3 > gpurun_out/profile_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python tools/profile_frame.py arch_1080p_128rpp 3 > gpurun_out/ncu_launches.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; tail -3 gpurun_out/bench.err
for f in gpurun_out/bench.json gpurun_out/bench_reference.json gpurun_out/bench_*rpp.json gpurun_out/bench_portal*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read()); r=d.get('roofline') or {}
    print(sys.argv[1].split('/')[-1], round(d['value'],3), 'Grays/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e'].get('ms_per_step',0),3), 'frac', round(r.get('frac',0),3), r.get('per_ray'), (d.get('cpu_baseline') or {}).get('value'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
