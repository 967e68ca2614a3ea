#!/bin/bash
# On an N-GPU box: bench.py at the given rank counts, then the cfg-4 / cfg-5 shapes on all GPUs.
# usage: scripts/gpu_scale.sh <tag> <ngpus...>
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
PORT=29520
for N in "$@"; do
  PORT=$((PORT+1))
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --no-secondary > $OUT/bench_n1.json 2> $OUT/bench_n1.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
  fi
  echo "N=$N rc=$?"; python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_n$N.json"))
    print("N=%d value %.0f periods/s  step %.3f ms  e2e %.0f periods/s (%.3f ms)  kernel %.3f ms" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms_per_launch"]))
except Exception as e:
    print("failed", e); print(open("$OUT/bench_n$N.err").read()[-800:])
PY
done
NMAX=${@: -1}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NMAX --master-addr 127.0.0.1 --master-port 29540 scripts/gpu_multi_features.py $OUT/features_n$NMAX.json --big 2>&1 | grep -E "batch_power|search_planets|Error|error" | cut -c1-600
