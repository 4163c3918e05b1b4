#!/bin/bash
# GPU box: parked light samples (VP_LIGHT_DEFER, the default build) against the variant build without them
P=$PWD/raytracer-voxpopuli_b200
run() { echo "$1 $(env $2 python tools/shard_probe.py ${3:-2 1} ${4:-60} 2>&1 | tail -1 | cut -c1-40)"; }
{
python -m pytest tests -m gpu -x -q 2>&1 | tail -1
for c in "2 1" "3 1" "4 1" "5 1" "1 1" "2 8"; do
run "defer   $c" "A=0" "$c"
run "nodefer $c" "VP_PRODUCT_SO=$P/libvoxpop_b200_nodefer.so" "$c"
done
} | tee gpurun_out/${1:-x4}_sweep.log
