#!/bin/bash
# A/B run of experimental builds (scripts/build_variants.py).  Usage: scripts/gpu_variants.sh <tag> [workloads...]
TAG=${1:-var}; shift
WLS=${@:-cfg1}
OUT=gpurun_out/$TAG; mkdir -p $OUT
for LIBF in tls_b200/variants/lib_*.so; do
  V=$(basename $LIBF .so); V=${V#lib_}
  for WL in $WLS; do
    EXTRA=""; [ "$WL" = "cfg2" ] && EXTRA="--max-periods 6000"
    TLSB200_LIB=$PWD/$LIBF python bench.py --workload $WL --steps 10 --warmup 3 --no-cpu-baseline --no-secondary $EXTRA > $OUT/bench_${V}_$WL.json 2> $OUT/bench_${V}_$WL.err
    python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_${V}_$WL.json"))
    print("%-10s %-12s kernel %.3f ms  step %.3f ms  e2e %.3f ms  frac %.3f" % ("$V", "$WL", d["roofline"]["kernel_ms_per_launch"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"]))
except Exception as e:
    print("$V $WL failed", e); print(open("$OUT/bench_${V}_$WL.err").read()[-600:])
PY
  done
done
