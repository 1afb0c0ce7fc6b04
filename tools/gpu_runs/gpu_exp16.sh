#!/bin/bash
# development: one slice per warp, after the code-size cut (coder) and the overlapped context (decoder)
mkdir -p gpurun_out
P="python bench.py --workload few --batch 1 --steps 2 --warmup 3 --no-cpu --no-e2e"
timeout 200 $P > gpurun_out/few_lone2.json 2> gpurun_out/few_lone2.err
python - gpurun_out/few_lone2.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], "value", round(d["value"],2), {k:round(v,2) for k,v in (d.get("kernel_ms_per_step") or {}).items() if k in ("symbolize","code","decode","pack_gather")}, {k:round(v,1) for k,v in d["decisions"].items() if "cycles" in k})
PY
