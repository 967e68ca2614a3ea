#!/bin/bash
# A/B of variant builds on cfg-2: kernel time (bench) and DRAM bytes per period (ncu, two metrics only).
OUT=gpurun_out/${1:-trab}; mkdir -p $OUT
for LIBF in tls_b200/variants/lib_*.so; do
  V=$(basename $LIBF .so); V=${V#lib_}
  TLSB200_LIB=$PWD/$LIBF python bench.py --workload cfg2 --max-periods 6000 --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/b_$V.json 2> $OUT/b_$V.err
  TLSB200_LIB=$PWD/$LIBF timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:tlsb_search -s 3 -c 1 --csv --log-file $OUT/n_$V.csv \
      python bench.py --workload cfg2 --max-periods 3000 --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > /dev/null 2>&1
  python - <<PY
import json, csv
d = json.load(open("$OUT/b_$V.json"))
rows = [r for r in csv.reader(open("$OUT/n_$V.csv")) if len(r) > 10 and r[0].isdigit()]
m = {r[-3]: (float(r[-1].replace(",", "")), r[-2]) for r in rows}
sc = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
rd = m["dram__bytes_read.sum"][0] * sc[m["dram__bytes_read.sum"][1]]; wr = m["dram__bytes_write.sum"][0] * sc[m["dram__bytes_write.sum"][1]]
print("%-8s kernel %.3f ms per 6000 periods   dram read %.2f + write %.2f = %.2f MB per period" % ("$V", d["roofline"]["kernel_ms_per_launch"], rd / 3000 / 1e6, wr / 3000 / 1e6, (rd + wr) / 3000 / 1e6))
PY
done
