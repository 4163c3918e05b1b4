#!/bin/bash
# knob sweep for small shards (one rank's share of an 8-GPU frame) on one GPU
python tools/shard_probe.py 2 1 40
python tools/shard_probe.py 2 8
for l in 5 6; do VP_LANES=$l python tools/shard_probe.py 2 8; done
for p in 2 4 16 32; do VP_PATHS_PER_THREAD=$p python tools/shard_probe.py 2 8; done
for m in 1 2 4 5; do VP_MIN_BLOCKS=$m python tools/shard_probe.py 2 8; done
for r in 12 16 24; do VP_REFILL=$r python tools/shard_probe.py 2 8; done
for b in 2 3 4; do VP_BLOCKS_SHADE=$b python tools/shard_probe.py 2 8; done
VP_GRAPH=0 python tools/shard_probe.py 2 8
VP_PREPARE=0 python tools/shard_probe.py 2 8
