for F in 0 25 35 45 55 65; do
  echo "TP_FRAC=$F"
  TLSB_TP_FRAC=$F python bench.py --workload cfg2 --max-periods 6000 --steps 3 --warmup 2 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('  cfg2 uniform %.3f ms  tiled_widths %s chunk %s' % (d['roofline']['kernel_ms_per_launch'], d['roofline']['layout'].get('tiled_widths'), d['roofline']['layout']['chunk']))"
  TLSB_TP_FRAC=$F python bench.py --workload cfg3 --steps 3 --warmup 2 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('  cfg3 uniform %.3f ms  tiled_widths %s chunk %s' % (d['roofline']['kernel_ms_per_launch'], d['roofline']['layout'].get('tiled_widths'), d['roofline']['layout']['chunk']))"
done
