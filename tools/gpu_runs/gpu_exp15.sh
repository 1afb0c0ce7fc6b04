#!/bin/bash
# development: the straight-line (one slice per warp) coders -- parity first, then the gain
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "every_form or encoder_packets or decoder_pictures or damaged or fate" > gpurun_out/pytest15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest15.log
tail -5 gpurun_out/pytest15.log
P="python bench.py --workload few --batch 1 --steps 2 --warmup 3 --no-cpu --no-e2e"
timeout 200 $P > gpurun_out/few_lone.json 2> gpurun_out/few_lone.err
FFGPU_LONE=0 timeout 200 $P > gpurun_out/few_warp.json 2> gpurun_out/few_warp.err
timeout 300 python bench.py --workload C4 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/c4_lone.json 2> gpurun_out/c4_lone.err
for f in gpurun_out/few_lone.json gpurun_out/few_warp.json gpurun_out/c4_lone.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], "value", round(d["value"],2), {k:round(v,2) for k,v in (d.get("kernel_ms_per_step") or {}).items() if k in ("symbolize","code","decode","pack_gather")}, {k:round(v,1) for k,v in d["decisions"].items() if "cycles" in k})
except Exception as ex:
    print(sys.argv[1], "ERR", ex)
PY
done
