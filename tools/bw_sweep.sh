#!/bin/bash
# knob sweep (GPU box): ms/frame of cfg2, value in Mrays/s
run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), 'ms', round(d['value']), 'Mrays/s  extend', round(d['roofline']['kernel_ms_per_step'],3), 'ms  e2e', round(d['e2e']['ms_per_step'],3))"; }
run "default"
for r in 16 18 22; do VP_REFILL=$r run "refill=$r"; done
VP_LANES=5 run "lanes=5"
