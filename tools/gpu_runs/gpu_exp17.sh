#!/bin/bash
# development: where the lone lane's cycles go in the straight-line coders
mkdir -p gpurun_out
P="python bench.py --workload few --batch 1 --steps 1 --warmup 1 --no-cpu --no-e2e"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_decode|k_code_range' -c 2 -o gpurun_out/prof17 $P > gpurun_out/ncu17.log 2>&1
tail -3 gpurun_out/ncu17.log
