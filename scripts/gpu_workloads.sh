#!/bin/bash
# Bench lines for the other workloads (parity-test configs; not the headline).  Usage: scripts/gpu_workloads.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for WL in cfg1_500ppm tutorial01 cfg3; do
  python bench.py --workload $WL --steps 5 --warmup 3 --cpu-seconds 6 > $OUT/bench_$WL.json 2> $OUT/bench_$WL.err; echo "$WL rc=$?"; cut -c1-250 $OUT/bench_$WL.json; tail -3 $OUT/bench_$WL.err
done
python bench.py --workload cfg2 --max-periods 6000 --steps 3 --warmup 3 --cpu-seconds 6 > $OUT/bench_cfg2.json 2> $OUT/bench_cfg2.err; echo "cfg2 rc=$?"; cut -c1-250 $OUT/bench_cfg2.json; tail -3 $OUT/bench_cfg2.err
