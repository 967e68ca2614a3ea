#!/bin/bash
# Final check of a build plus an A/B against variant builds (tls_b200/variants/lib_*.so) on cfg-1.
# Usage: scripts/gpu_final_ab.sh <tag>
TAG=${1:-gate}
OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_gpu.log
python bench.py > $OUT/bench.json 2> $OUT/bench.err
python -c "
import json; d=json.load(open('$OUT/bench.json')); print('value %.0f e2e %.0f frac %.3f kernel %.3f ms launches %d' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch'], d['gpu_launches']), d['parity'], d['secondary']['cfg2']['value'], d['secondary']['power']['wall_s'], d['cpu_baseline']['value'])"
run() {  # name lib workloads...
  local NAME=$1 LIB=$2; shift 2
  for WL in "$@"; do
    TLSB200_LIB=$LIB python bench.py --workload $WL --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/bench_${NAME}_$WL.json 2> $OUT/bench_${NAME}_$WL.err
    python -c "
import json; d = json.load(open('$OUT/bench_${NAME}_$WL.json')); print('%-10s %-12s kernel %.3f ms  step %.3f ms  frac %.3f' % ('$NAME', '$WL', d['roofline']['kernel_ms_per_launch'], d['ms_per_step'], d['roofline']['frac']))"
  done
}
run main $PWD/tls_b200/libtlsb200.so cfg1
shopt -s nullglob
for LIBF in tls_b200/variants/lib_*.so; do
  V=$(basename $LIBF .so); V=${V#lib_}
  run $V $PWD/$LIBF cfg1
done
