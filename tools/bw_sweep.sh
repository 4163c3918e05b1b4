#!/bin/bash
# compile-time variants (built into alt6/<name>/ beforehand) against the shipped library (GPU box)
run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), 'ms', round(d['value']), 'Mrays/s  extend', round(d['roofline']['kernel_ms_per_step'],3), 'ms  e2e', round(d['e2e']['ms_per_step'],3))"; }
run "shipped"
cp raytracer-voxpopuli_b200/libvoxpop_b200.so /tmp/keep.so
for v in alt6/*/; do
  n=$(basename $v)
  cp $v/libvoxpop_b200.so raytracer-voxpopuli_b200/libvoxpop_b200.so; run "variant $n"
done
cp /tmp/keep.so raytracer-voxpopuli_b200/libvoxpop_b200.so
run "shipped (again)"
