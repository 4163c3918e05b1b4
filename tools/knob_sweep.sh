#!/bin/bash
# GPU box: the defaults (six passes in flight, extend cap 2) and their neighbourhood
run() { echo "$1 $(env $1 python tools/shard_probe.py ${2:-2 1} ${3:-60} 2>&1 | tail -1 | cut -c1-40)"; }
{
python -m pytest tests -m gpu -x -q 2>&1 | tail -1
run "A=0"
run "VP_OV_SHADE=1"
run "VP_OV_SHADOW=1"
run "VP_OV_EXTEND=1"
run "VP_LANES=8"
run "A=0" "3 1"
run "A=0" "4 1"
run "A=0" "5 1"
run "A=0" "1 1"
run "A=0" "2 2"
run "A=0" "2 4"
run "A=0" "2 8" 120
run "VP_SHARD_LANES=0" "2 8" 120
} | tee gpurun_out/${1:-x4}_sweep.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${1:-x4}_n1.json 2> gpurun_out/${1:-x4}_n1.err; python -c "
import json; d=json.loads(open('gpurun_out/${1:-x4}_n1.json').read().strip().splitlines()[-1]); print('bench', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], 'sync', d['e2e']['synchronous']['ms_per_step'], 'hbm', d['detail']['hbm_in_use'], 'lanes', d['detail']['frames_in_flight'])"
