#!/bin/bash
# Bench lines (kernel ms, parity) for a list of workloads.  Usage: scripts/gpu_bench_all.sh <tag> [workloads...]
TAG=${1:-ball}; shift
WLS=${@:-cfg1 cfg1_500ppm tutorial01 cfg3 cfg2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
for WL in $WLS; do
  EXTRA=""; [ "$WL" = "cfg2" ] && EXTRA="--max-periods 6000"
  timeout 900 python bench.py --workload $WL --steps 5 --warmup 3 --cpu-seconds 1 --no-secondary $EXTRA > $OUT/bench_$WL.json 2> $OUT/bench_$WL.err
  python -c "
import json
try:
    d = json.load(open('$OUT/bench_$WL.json')); l = d['roofline']['layout']
    print('%-12s kernel %.3f ms  value %.4g  e2e %.4g  frac %.3f  %s threads %d x %d  R %d  parity %s %.1e' % ('$WL', d['roofline']['kernel_ms_per_launch'], d['value'], d['e2e']['value'], d['roofline']['frac'], l.get('path'), l['threads'], l['ctas_per_sm'], l['block'], d['parity']['rows_equal'], d['parity']['chi2_max_rel_err']))
except Exception as e:
    print('$WL failed', e); print(open('$OUT/bench_$WL.err').read()[-800:])"
done
