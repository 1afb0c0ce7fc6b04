#!/bin/bash
# development: default workload still at its level with the final build; the lone coders on `few`
mkdir -p gpurun_out
timeout 70 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/c2_final_quick.json 2> gpurun_out/c2_final_quick.err
timeout 40 python bench.py --workload few --batch 1 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/few_lone3.json 2> gpurun_out/few_lone3.err
for f in gpurun_out/c2_final_quick.json gpurun_out/few_lone3.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], "value", round(d["value"],2), {k:round(v,2) for k,v in (d.get("kernel_ms_per_step") or {}).items() if k in ("symbolize","code","decode","pack_gather")}, {k:round(v,1) for k,v in d["decisions"].items() if "cycles" in k})
except Exception as ex:
    print(sys.argv[1], "ERR", ex)
PY
done
