#!/bin/bash
# GPU box: A/B loop over the library's environment knobs - steady-state ms per frame (or frame share) of a config.
# usage: tools/knob_sweep.sh <tag> ["<cfg> <world>"] ; edit the list below; results in gpurun_out/<tag>_sweep.log.
# (How the shipped defaults were found: VP_OV_* 4/2/2 -> 2/1/1, VP_LANES 4 -> 6, VP_SSPARE_MIN, VP_BANDS / VP_BAND_STAGGER;
#  variant builds - build.build_variant(name, ["MACRO=..."]) - are selected with VP_PRODUCT_SO=<path>.)
what=${2:-2 1}
run() { echo "$1 $(env $1 python tools/shard_probe.py $what 60 2>&1 | tail -1 | cut -c1-40)"; }
{
run "A=0"
for e in 1 2 3 4; do run "VP_OV_EXTEND=$e"; done
for e in 1 2 3; do run "VP_OV_SHADOW=$e"; run "VP_OV_SHADE=$e"; done
for l in 4 6 8; do run "VP_LANES=$l"; done
run "VP_SSPARE_MIN=524288"
run "VP_SPARE=0"
run "VP_REFILL_SPARE=20"
run "VP_REFILL_SPARE=28"
} | tee gpurun_out/${1:-sweep}_sweep.log
