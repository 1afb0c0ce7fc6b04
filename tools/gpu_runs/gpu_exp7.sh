#!/bin/bash
# development: heavy/light coder schedule A/B
set -x
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e"
run() { name=$1; shift; env "$@" $B $EXTRA > gpurun_out/b7_$name.json 2>gpurun_out/b7_$name.err || tail -3 gpurun_out/b7_$name.err; }
run off FFGPU_HEAVY_STRIDE=0
run s4_f200 FFGPU_HEAVY_STRIDE=4
run s4_f150 FFGPU_HEAVY_STRIDE=4 FFGPU_HEAVY_FACTOR=150
run s4_f300 FFGPU_HEAVY_STRIDE=4 FFGPU_HEAVY_FACTOR=300
run s8_f200 FFGPU_HEAVY_STRIDE=8
run s2_f150 FFGPU_HEAVY_STRIDE=2 FFGPU_HEAVY_FACTOR=150
run s4_generic FFGPU_HEAVY_STRIDE=4 FFGPU_DEC_GENERIC=1
run s8_f150 FFGPU_HEAVY_STRIDE=8 FFGPU_HEAVY_FACTOR=150
EXTRA=--synth run synth_off FFGPU_HEAVY_STRIDE=0
EXTRA=--synth run synth_s4 FFGPU_HEAVY_STRIDE=4
EXTRA="--source mandelbrot" run mandel_off FFGPU_HEAVY_STRIDE=0
EXTRA="--source mandelbrot" run mandel_s4 FFGPU_HEAVY_STRIDE=4
for f in gpurun_out/b7_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d["value"],1), {k:round(v,2) for k,v in (d.get("kernel_ms_per_step") or {}).items() if k in ("code","decode","sort")})
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
