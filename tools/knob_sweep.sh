#!/bin/bash
# GPU box: grid caps while passes overlap for multi-volume scenes (VP_OV_MULTI)
run() { echo "$1 $(env $1 python tools/shard_probe.py ${2:-4 1} ${3:-60} 2>&1 | tail -1 | cut -c1-40)"; }
{
run "A=0"
run "VP_OV_MULTI=1 VP_OV_EXTEND=4 VP_OV_SHADOW=3 VP_OV_SHADE=2"
run "VP_OV_MULTI=1 VP_OV_EXTEND=5 VP_OV_SHADOW=4 VP_OV_SHADE=2"
run "VP_OV_MULTI=1 VP_OV_EXTEND=0 VP_OV_SHADOW=0 VP_OV_SHADE=2"
run "VP_OV_MULTI=1 VP_OV_EXTEND=3 VP_OV_SHADOW=2 VP_OV_SHADE=1"
run "A=0" "5 1"
run "VP_OV_MULTI=1 VP_OV_EXTEND=4 VP_OV_SHADOW=3 VP_OV_SHADE=2" "5 1"
run "VP_OV_MULTI=1 VP_OV_EXTEND=0 VP_OV_SHADOW=0 VP_OV_SHADE=2" "5 1"
run "VP_OV_MULTI=1 VP_OV_EXTEND=3 VP_OV_SHADOW=2 VP_OV_SHADE=1" "5 1"
} | tee gpurun_out/${1:-x4}_sweep.log
