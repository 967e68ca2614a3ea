#!/bin/bash
# Next experiment: more, thinner warps in the resident kernel.  TLSB_THREADS=320/384 launches two CTAs
# of 10/12 warps per SM at 96/80 registers; TLSB_BLOCK picks R.  Every run carries bench.py's parity spot check
# against the oracle (64 periods), so a broken variant shows up as rows_equal False.
# Usage: scripts/gpu_threads_ab.sh <tag> [workloads...]      (default: cfg1 cfg1_500ppm)
TAG=${1:-thr}; shift
WLS=${@:-cfg1 cfg1_500ppm}
OUT=gpurun_out/$TAG; mkdir -p $OUT
run() {  # name env...
  local NAME=$1; shift
  for WL in $WLS; do
    env "$@" python bench.py --workload $WL --steps 10 --warmup 3 --cpu-seconds 1 --no-secondary > $OUT/bench_${NAME}_$WL.json 2> $OUT/bench_${NAME}_$WL.err
    python -c "
import json
try:
    d = json.load(open('$OUT/bench_${NAME}_$WL.json')); l = d['roofline']['layout']
    print('%-10s %-12s kernel %.3f ms  frac %.3f  threads %d x %d  R %d  parity %s %.1e' % ('$NAME', '$WL', d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'], l['threads'], l['ctas_per_sm'], l['block'], d['parity']['rows_equal'], d['parity']['chi2_max_rel_err']))
except Exception as e:
    print('$NAME $WL failed', e); print(open('$OUT/bench_${NAME}_$WL.err').read()[-500:])"
  done
}
run base    TLSB_NOP=1
run t256r5  TLSB_BLOCK=5
run t320r5  TLSB_THREADS=320 TLSB_BLOCK=5
run t320r7  TLSB_THREADS=320
run t384r5  TLSB_THREADS=384 TLSB_BLOCK=5
run t384r7  TLSB_THREADS=384
