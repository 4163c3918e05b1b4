#!/bin/bash
# GPU box: one traversal stage per bounce (VP_MERGE) against the separate extend / shadow stages
run() { echo "$1 $(env $1 python tools/shard_probe.py ${2:-2 1} ${3:-60} 2>&1 | tail -1 | cut -c1-40)"; }
{
python -m pytest tests -m gpu -x -q 2>&1 | tail -1
run "VP_MERGE=0"
run "VP_MERGE=1"
run "VP_MERGE=1 VP_OV_EXTEND=5"
run "VP_MERGE=1 VP_OV_EXTEND=6"
run "VP_MERGE=1 VP_OV_EXTEND=3"
run "VP_MERGE=0" "2 8" 120
run "VP_MERGE=1" "2 8" 120
run "VP_MERGE=0" "2 2"
run "VP_MERGE=1" "2 2"
run "VP_MERGE=0" "3 1"
run "VP_MERGE=1" "3 1"
run "VP_MERGE=0" "1 1"
run "VP_MERGE=1" "1 1"
} | tee gpurun_out/${1:-x4}_sweep.log
