#!/bin/bash
# tools/sweep.sh -- run bench.py over several (L, threads) layouts on the GPU box; prints one line each.
for cfg in "$@"; do
  set -- $cfg
  python bench.py --steps 2 --warmup 1 --no-cpu --L $1 --threads $2 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('L=%d NT=%d occ=%d regs=%d  %.3e pstep/s  %.1f ms' % (d['layout']['scan_items_per_lane'], d['layout']['threads_per_filter'], d['layout']['filters_per_sm'], d['layout']['registers_per_thread'], d['value'], d['ms_per_step']))
except Exception as e: print('failed', e)
"
done
