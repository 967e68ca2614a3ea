#!/bin/bash
# fp32 filter pass: parity (filter on == filter off bit for bit, goldens) and the A/B of the kernel time.
# Usage: scripts/gpu_filter_check.sh <tag> [workloads...]
TAG=${1:-filt}; shift
WLS=${@:-cfg1 cfg1_500ppm}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_filter.py tests/test_gpu_parity.py -x -q 2>&1 | tail -15 > $OUT/pytest.log
cat $OUT/pytest.log
python scripts/gpu_filter_stats.py $WLS 2>&1 | tee $OUT/stats.log
for WL in $WLS; do
  for F in 1 0; do
    [ $F = 0 ] && [ "$WL" != "cfg1" ] && continue
    TLSB_FILTER=$F timeout 600 python bench.py --workload $WL --steps 10 --warmup 3 --cpu-seconds 1 --no-secondary > $OUT/bench_f${F}_$WL.json 2> $OUT/bench_f${F}_$WL.err
    python -c "
import json
try:
    d = json.load(open('$OUT/bench_f${F}_$WL.json')); l = d['roofline']['layout']
    print('filter=$F %-12s kernel %.3f ms  frac %.3f  threads %d x %d  R %d  parity %s %.1e' % ('$WL', d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'], l['threads'], l['ctas_per_sm'], l['block'], d['parity']['rows_equal'], d['parity']['chi2_max_rel_err']))
except Exception as e:
    print('filter=$F $WL failed', e); print(open('$OUT/bench_f${F}_$WL.err').read()[-800:])"
  done
done
