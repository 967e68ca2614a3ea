#!/bin/bash
# A/B of the tiled layout: one 512-thread CTA per SM vs two 256-thread CTAs (TLSB_TILED), cfg-2 and cfg-3.
OUT=gpurun_out/${1:-tab}; mkdir -p $OUT
for T in 512 256; do
  for WL in cfg2 cfg3; do
    EXTRA=""; [ "$WL" = "cfg2" ] && EXTRA="--max-periods 6000"
    TLSB_TILED=$T timeout 600 python bench.py --workload $WL --steps 3 --warmup 3 --cpu-seconds 1 --no-secondary $EXTRA > $OUT/b_${WL}_$T.json 2> $OUT/b_${WL}_$T.err
    python -c "
import json
try:
    d = json.load(open('$OUT/b_${WL}_$T.json')); l = d['roofline']['layout']
    print('TILED=$T %-6s kernel %.3f ms  threads %d x %d chunk %d tiled_widths %d parity %s' % ('$WL', d['roofline']['kernel_ms_per_launch'], l['threads'], l['ctas_per_sm'], l['chunk'], l['tiled_widths'], d['parity']['rows_equal']))
except Exception as e:
    print('$WL $T failed', e); print(open('$OUT/b_${WL}_$T.err').read()[-600:])"
  done
done
