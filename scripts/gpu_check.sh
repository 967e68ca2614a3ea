#!/bin/bash
# Runs on the B200 box (gpurun): GPU parity tests, bench, ncu launch list + one full capture.
# Usage: scripts/gpu_check.sh <tag> [workload]
set -u
TAG=${1:-r01}
WL=${2:-cfg1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/nvidia_smi.csv 2>&1
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
python bench.py --workload $WL > $OUT/bench_$WL.json 2> $OUT/bench_$WL.err; echo "bench rc=$?"
cat $OUT/bench_$WL.json; tail -5 $OUT/bench_$WL.err
python bench.py --impl reference --steps 3 --warmup 1 --workload $WL > $OUT/bench_ref_$WL.json 2>> $OUT/bench_$WL.err
cat $OUT/bench_ref_$WL.json
# launch list (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_$WL.csv \
    python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/ncu_launches.log 2>&1
# one full capture of the search kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tlsb_search -s 3 -c 1 -f -o $OUT/prof_$WL \
    python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/ncu_full.log 2>&1
ls -la $OUT
