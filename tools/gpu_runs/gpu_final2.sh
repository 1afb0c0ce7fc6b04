#!/bin/bash
# round 2, last run: the few-slice workloads with the straight-line coders (each bench line
# gates on packet parity with the CPU reference and on the round trip), then the GPU tests
# that force every form of the slice coders
set -x
mkdir -p gpurun_out
for wl in C4 C5 C3; do
  timeout 125 python bench.py --workload $wl --steps 2 --warmup 3 --e2e-repeat 1 --e2e-depth 3 --no-pageable > gpurun_out/r02_bench_${wl}_lone.json 2> gpurun_out/${wl}_lone.err; echo "$wl rc=$?"; tail -2 gpurun_out/${wl}_lone.err
done
for f in gpurun_out/r02_bench_*_lone.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    e=d.get("e2e") or {}
    print(sys.argv[1], "value", round(d["value"],1), "e2e", e.get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"), {k:round(v,2) for k,v in (d.get("kernel_ms_per_step") or {}).items() if k in ("symbolize","code","decode")})
except Exception as ex:
    print(sys.argv[1], "ERR", ex)
PY
done
timeout 110 python -m pytest tests -m gpu -q -x -k "every_form or fate or damaged or wider or resident or (decoder_pictures and (bgr0 or gbrp16le or rgb48le)) or (version4_rgb and bgra)" > gpurun_out/pytest_final2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final2.log
tail -4 gpurun_out/pytest_final2.log
