#!/bin/bash
# One ncu --set full capture of the search kernel per workload (source-level).  Usage: scripts/gpu_prof.sh <tag> [workloads...]
TAG=${1:-prof}; shift
WLS=${@:-cfg1}
OUT=gpurun_out/$TAG; mkdir -p $OUT
for WL in $WLS; do
  EXTRA=""; [ "$WL" = "cfg2" ] && EXTRA="--max-periods 3000"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:tlsb_search -s 3 -c 1 -f -o $OUT/prof_$WL \
      python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --no-secondary $EXTRA > $OUT/ncu_$WL.log 2>&1
  tail -3 $OUT/ncu_$WL.log
done
ls -la $OUT
