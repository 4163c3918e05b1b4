#!/bin/bash
# brick-walk A/B (GPU box): ms/frame of cfg2, value in Mrays/s
run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), 'ms', round(d['value']), 'Mrays/s  extend', round(d['roofline']['kernel_ms_per_step'],3), 'ms')"; }
VP_BRICKWALK=0 run "cell walk (old)"
run "brick walk 8 blocks/SM"
for r in 16 24 28; do VP_REFILL=$r run "refill=$r"; done
if [ -f alt6/libvoxpop_b200.so ]; then
  cp raytracer-voxpopuli_b200/libvoxpop_b200.so /tmp/keep.so; cp alt6/libvoxpop_b200.so raytracer-voxpopuli_b200/libvoxpop_b200.so
  run "brick walk 6 blocks/SM"
  VP_BRICKWALK=0 run "cell walk 6 blocks/SM"
  cp /tmp/keep.so raytracer-voxpopuli_b200/libvoxpop_b200.so
fi
