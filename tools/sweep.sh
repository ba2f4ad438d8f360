#!/bin/bash
# tools/sweep.sh -- run bench.py over several "L threads [proposals]" layouts on the GPU box; one line each.
for cfg in "$@"; do
  set -- $cfg
  P=${3:-4096}
  python bench.py --steps 2 --warmup 1 --no-cpu --no-pmmh --L $1 --threads $2 --proposals $P 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('P=%d L=%d NT=%d occ=%d regs=%d  %.3e pstep/s  %.1f ms' % (d['config']['proposals_per_gpu'], d['layout']['scan_items_per_lane'], d['layout']['threads_per_filter'], d['layout']['filters_per_sm'], d['layout']['registers_per_thread'], d['value'], d['ms_per_step']))
except Exception as e: print('failed', e)
"
done
