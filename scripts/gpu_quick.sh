#!/bin/bash
# Quick GPU check: parity tests + one bench line per workload.  Usage: scripts/gpu_quick.sh <tag> [pytest-args]
TAG=${1:-q}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
for WL in cfg1 cfg1_500ppm cfg3; do
  python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_$WL.json 2> $OUT/bench_$WL.err
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$WL.json"))
    print("$WL", "value %.0f periods/s  %.3f ms/step  kernel %.3f ms  frac %.3f  path %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["frac"], d["roofline"]["path"][:8]))
except Exception as e:
    print("$WL failed", e); print(open("$OUT/bench_$WL.err").read()[-800:])
PY
done
python bench.py --workload cfg2 --max-periods 6000 --steps 3 --warmup 2 --no-cpu-baseline > $OUT/bench_cfg2.json 2> $OUT/bench_cfg2.err
python -c "
import json; d = json.load(open('$OUT/bench_cfg2.json')); print('cfg2', 'value %.0f periods/s  %.3f ms/step frac %.3f path %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['path'][:8]))"
