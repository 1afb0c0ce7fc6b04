#!/bin/bash
# development: full parity with the bulk-copy stage A and the heavy/light schedule; A/B of both
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest8.log
tail -8 gpurun_out/pytest8.log
B="python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/b8_$name.json 2>gpurun_out/b8_$name.err || tail -3 gpurun_out/b8_$name.err; }
run default X=1
run legacyA FFGPU_STAGE_A=legacy
EXTRA=--synth run synth X=1
EXTRA=--synth run synth_legacyA FFGPU_STAGE_A=legacy
EXTRA="--workload C1" run C1 X=1
EXTRA="--workload C1" run C1_legacyA FFGPU_STAGE_A=legacy
EXTRA="--workload C5 --batch 24" run C5 X=1
EXTRA="--workload C5 --batch 24" run C5_legacyA FFGPU_STAGE_A=legacy
for f in gpurun_out/b8_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d["value"],1), {k:round(v,2) for k,v in (d.get("kernel_ms_per_step") or {}).items() if k in ("symbolize","code","decode")})
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
