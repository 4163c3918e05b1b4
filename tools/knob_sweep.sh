#!/bin/bash
# GPU box: steady-state frame time against the resident blocks per SM of the stages when passes overlap (single-volume scenes)
run() { echo "$1 $(env $1 python tools/shard_probe.py ${2:-2 1} 40 2>&1 | tail -1 | cut -c1-40)"; }
{
run "A=0"
run "VP_OV_EXTEND=0 VP_OV_SHADOW=0 VP_OV_SHADE=0"
run "VP_OV_EXTEND=4"
run "VP_OV_EXTEND=3"
run "VP_OV_EXTEND=4 VP_OV_SHADOW=1"
run "VP_OV_EXTEND=4 VP_OV_SHADE=1"
run "VP_OV_EXTEND=4 VP_OV_SHADE=3"
run "VP_OV_EXTEND=4 VP_OV_SHADOW=3"
run "VP_OV_EXTEND=3 VP_OV_SHADOW=1 VP_OV_SHADE=1"
run "VP_LANES=5"
run "VP_LANES=6"
run "VP_OV_EXTEND=4 VP_LANES=6"
run "VP_OV_EXTEND=6 VP_OV_SHADOW=3 VP_OV_SHADE=3 VP_LANES=6"
run "A=0" "3 1"
run "VP_OV_EXTEND=4" "3 1"
run "A=0" "1 1"
run "A=0" "2 8"
run "A=0" "2 2"
run "VP_OV_EXTEND=0 VP_OV_SHADOW=0 VP_OV_SHADE=0" "2 2"
} | tee gpurun_out/${1:-x4}_sweep.log
