#!/bin/bash
# usage: scripts/gpu_multi.sh <tag> <ngpus...>   e.g. scripts/gpu_multi.sh r01g 1 2
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
for N in "$@"; do
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 > $OUT/bench_n1.json 2> $OUT/bench_n1.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
  fi
  echo "N=$N rc=$?"; cut -c1-200 $OUT/bench_n$N.json; tail -3 $OUT/bench_n$N.err
done
