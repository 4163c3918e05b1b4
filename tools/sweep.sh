#!/bin/bash
# tuning sweep helper (GPU box): prints ms/frame of cfg2 for a few knob settings
for r in 12 16 20 22 26 30; do
  VP_REFILL=$r python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('refill $r', round(d['ms_per_step'],3))"
done
for b in 3 4 5 6; do
  VP_BLOCKS_EXTEND=$b VP_BLOCKS_SHADOW=$b python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('blocks/SM $b', round(d['ms_per_step'],3))"
done
