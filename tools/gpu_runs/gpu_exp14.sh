#!/bin/bash
# development: where a lone lane's cycles go (few-slice launch, one slice per warp)
set -x
mkdir -p gpurun_out
P="python bench.py --workload few --batch 1 --steps 1 --warmup 1 --no-cpu --no-e2e"
timeout 300 $P > gpurun_out/plain14.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_decode|k_code_records|k_chain_states' -c 3 -o gpurun_out/prof14 $P > gpurun_out/ncu14.log 2>&1
tail -c 600 gpurun_out/plain14.log
tail -5 gpurun_out/ncu14.log
