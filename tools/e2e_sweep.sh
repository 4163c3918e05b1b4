#!/bin/bash
# end-to-end (vp_render, host buffers): staggered band starts (GPU box)
run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1  device', round(d['ms_per_step'],3), 'ms', round(d['value']), 'Mrays/s   e2e', round(d['e2e']['ms_per_step'],3), 'ms', round(d['e2e']['value']), 'Mrays/s')"; }
VP_BAND_STAGGER=-1 run "bands=4 no stagger"
for g in 0 1 2 3 5; do VP_BAND_STAGGER=$g run "bands=4 stagger=$g"; done
for g in 0 1 2; do VP_BANDS=8 VP_BAND_STAGGER=$g run "bands=8 stagger=$g"; done
VP_BANDS=6 VP_BAND_STAGGER=1 run "bands=6 stagger=1"
