#!/bin/bash
# experiment: cell-kernel shapes (threads, look-ahead) on C4 / C3 / C4s8
for shape in 11 12 13 14 15 16 17; do
  for w in C4 C3 C4s8; do
    CDL_SELL_SHAPE=$shape python bench.py --workload $w --steps 20 --warmup 5 --no-e2e --no-cpu --no-parity 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('shape $shape $w value %.1f cells_ms %.4f frac %.3f full_mode %.1f' % (d['value'], r['ms_per_launch'], r['frac'], d['full_pass_mode']['value']))"
  done
done
