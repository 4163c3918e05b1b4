#!/bin/bash
# GPU box: parity suite, then steady-state frame times of all configs and of a 1/8 frame share.  usage: tools/probe_all.sh <tag>
tag=${1:-x3}
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest: $(tail -1 gpurun_out/${tag}_pytest.log)"
run() { echo "$1 $(env $2 python tools/shard_probe.py ${3:-2 1} 60 2>&1 | tail -1 | cut -c1-60)"; }
{
run cfg2 "A=1"
run cfg1 "A=1" "1 1"
run cfg3 "A=1" "3 1"
run cfg4 "A=1" "4 1"
run cfg5 "A=1" "5 1"
run cfg2-1/8 "A=1" "2 8"
} | tee gpurun_out/${tag}_probe.log
