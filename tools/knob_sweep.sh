#!/bin/bash
# GPU box: grid caps of the shadow / shade stages with the prepared shadow stage on for frame shares
run() { echo "$1 $(env $1 python tools/shard_probe.py ${2:-2 8} ${3:-120} 2>&1 | tail -1 | cut -c1-40)"; }
{
run "A=0"
run "VP_OV_SHADOW=1"
run "VP_OV_SHADOW=1 VP_OV_SHADE=1"
run "A=0" "2 1" 60
run "VP_OV_SHADOW=1" "2 1" 60
run "VP_OV_SHADOW=1 VP_OV_SHADE=1" "2 1" 60
run "VP_OV_SHADOW=1 VP_OV_SHADE=1" "3 1" 60
run "VP_OV_SHADOW=1 VP_OV_SHADE=1" "2 4"
} | tee gpurun_out/${1:-x4}_sweep.log
