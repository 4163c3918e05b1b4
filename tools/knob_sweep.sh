#!/bin/bash
# GPU box: variant builds against the default build, steady-state frame times
P=$PWD/raytracer-voxpopuli_b200
run() { echo "$1 $(env $2 python tools/shard_probe.py ${3:-2 1} 60 2>&1 | tail -1 | cut -c1-40)"; }
{
python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k sincos 2>&1 | tail -1
run default "A=0"
run u3 "VP_PRODUCT_SO=$P/libvoxpop_b200_u3.so"
run sh4 "VP_PRODUCT_SO=$P/libvoxpop_b200_sh4.so"
run default-cfg3 "A=0" "3 1"
run u3-cfg3 "VP_PRODUCT_SO=$P/libvoxpop_b200_u3.so" "3 1"
run sh4-cfg3 "VP_PRODUCT_SO=$P/libvoxpop_b200_sh4.so" "3 1"
run default-cfg4 "A=0" "4 1"
run mv7-cfg4 "VP_PRODUCT_SO=$P/libvoxpop_b200_mv7.so" "4 1"
run u3-cfg4 "VP_PRODUCT_SO=$P/libvoxpop_b200_u3.so" "4 1"
run default-cfg5 "A=0" "5 1"
run mv7-cfg5 "VP_PRODUCT_SO=$P/libvoxpop_b200_mv7.so" "5 1"
run u3-cfg5 "VP_PRODUCT_SO=$P/libvoxpop_b200_u3.so" "5 1"
run sh4-cfg5 "VP_PRODUCT_SO=$P/libvoxpop_b200_sh4.so" "5 1"
run u3-1/8 "VP_PRODUCT_SO=$P/libvoxpop_b200_u3.so" "2 8"
} | tee gpurun_out/${1:-x4}_sweep.log
