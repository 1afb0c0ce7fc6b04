#!/bin/bash
# development: stage B in two halves for few-slice streams -- parity and effect
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "encoder_packets or version4_rgb or full_size or two_pass or several or fate or pipelined" > gpurun_out/pytest12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest12.log
tail -6 gpurun_out/pytest12.log
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e"
$B --workload C5 > gpurun_out/b12_C5_split.json 2> gpurun_out/b12_C5_split.err; tail -2 gpurun_out/b12_C5_split.err
FFGPU_SPLIT=0 $B --workload C5 > gpurun_out/b12_C5_fused.json 2>/dev/null
$B --workload C3 > gpurun_out/b12_C3_split.json 2> gpurun_out/b12_C3_split.err; tail -2 gpurun_out/b12_C3_split.err
$B --workload C5 --batch 96 > gpurun_out/b12_C5_split_b96.json 2> gpurun_out/b12_C5_b96.err; tail -2 gpurun_out/b12_C5_b96.err
python bench.py --steps 3 --warmup 3 --no-cpu --source mandelbrot --e2e-repeat 3 > gpurun_out/b12_mandel_e2e.json 2>/dev/null
for f in gpurun_out/b12_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    e=d.get("e2e") or {}
    print(sys.argv[1], round(d["value"],1), "enc", d.get("encode_fps"), "dec", d.get("decode_fps"), {k:round(v,2) for k,v in (d.get("kernel_ms_per_step") or {}).items() if k in ("symbolize","code","decode")}, "e2e", e.get("value"), "pg", (e.get("pageable") or {}).get("value"))
except Exception as ex:
    print(sys.argv[1], "ERR", ex)
PY
done
