#!/bin/bash
# end-to-end (vp_render, host buffers) with different numbers of row bands and lane priorities (GPU box)
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1  device', round(d['ms_per_step'],3), 'ms   e2e', round(d['e2e']['ms_per_step'],3), 'ms', round(d['e2e']['value']), 'Mrays/s')"; }
for b in 1 4; do VP_BANDS=$b run "bands=$b prio=0"; done
for b in 2 3 4 6 8; do VP_LANE_PRIO=1 VP_BANDS=$b run "bands=$b prio=1"; done
