#!/bin/bash
# GPU box: hardware queues (CUDA_DEVICE_MAX_CONNECTIONS) against the passes in flight, whole frames and frame shares
run() { echo "$1 $(env $1 python tools/shard_probe.py ${2:-2 8} 120 2>&1 | tail -1 | cut -c1-40)"; }
{
for c in 16 32; do
run "CUDA_DEVICE_MAX_CONNECTIONS=$c"
run "CUDA_DEVICE_MAX_CONNECTIONS=$c VP_SHARD_LANES=4"
run "CUDA_DEVICE_MAX_CONNECTIONS=$c VP_SHARD_LANES=6"
run "CUDA_DEVICE_MAX_CONNECTIONS=$c VP_SHARD_LANES=8"
done
run "CUDA_DEVICE_MAX_CONNECTIONS=32" "2 1"
run "CUDA_DEVICE_MAX_CONNECTIONS=32 VP_LANES=6" "2 1"
run "CUDA_DEVICE_MAX_CONNECTIONS=32 VP_LANES=8" "2 1"
run "CUDA_DEVICE_MAX_CONNECTIONS=32 VP_SHARD_LANES=8" "2 4"
run "CUDA_DEVICE_MAX_CONNECTIONS=32 VP_SHARD_LANES=8" "2 2"
run "CUDA_DEVICE_MAX_CONNECTIONS=32" "5 1"
} | tee gpurun_out/${1:-x4}_sweep.log
