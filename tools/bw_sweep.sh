#!/bin/bash
# brick-walk knob sweep (GPU box): ms/frame of cfg2, value in Mrays/s
run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), 'ms', round(d['value']), 'Mrays/s  extend', round(d['roofline']['kernel_ms_per_step'],3), 'ms')"; }
run "brick walk unroll 2"
for r in 14 16 18 22 24; do VP_REFILL=$r run "refill=$r"; done
cp raytracer-voxpopuli_b200/libvoxpop_b200.so /tmp/keep.so
for u in 1 3 4; do
  if [ -f alt6/u$u/libvoxpop_b200.so ]; then cp alt6/u$u/libvoxpop_b200.so raytracer-voxpopuli_b200/libvoxpop_b200.so; run "unroll $u"; fi
done
cp /tmp/keep.so raytracer-voxpopuli_b200/libvoxpop_b200.so
