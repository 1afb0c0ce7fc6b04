#!/bin/bash
# development: A/B of decode-kernel variants + source-level profiles of both coder kernels
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
B="python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e"
L=$PWD/ffmpeg_ffv2_b200/build
$B > gpurun_out/b2_new.json 2> gpurun_out/b2_new.err; tail -2 gpurun_out/b2_new.err
FFGPU_DEC_GENERIC=1 $B > gpurun_out/b2_generic.json 2>/dev/null
FFGPU_LIB=$L/libffgpu_mb6.so $B > gpurun_out/b2_mb6.json 2>/dev/null
FFGPU_LIB=$L/libffgpu_mb8.so $B > gpurun_out/b2_mb8.json 2>/dev/null
FFGPU_LIB=$L/libffgpu_mb6a2.so $B > gpurun_out/b2_mb6a2.json 2>/dev/null
FFGPU_LIB=$L/libffgpu_mb6.so FFGPU_GATE_WAIT=12 $B > gpurun_out/b2_mb6_w12.json 2>/dev/null
FFGPU_LIB=$L/libffgpu_mb6.so $B --batch 256 > gpurun_out/b2_mb6_b256.json 2>/dev/null
for f in gpurun_out/b2_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d["value"]), {k:round(v,2) for k,v in d["kernel_ms_per_step"].items() if k in ("symbolize","code","decode","init_state","fill_state")})
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
$B --steps 1 --warmup 1 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_decode -s 2 -c 1 -o gpurun_out/prof2_dec $B --steps 1 --warmup 1 > gpurun_out/ncu_dec.log 2>&1
echo "ncu dec rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_code_range -s 2 -c 1 -o gpurun_out/prof2_code $B --steps 1 --warmup 1 > gpurun_out/ncu_code.log 2>&1
echo "ncu code rc=$?"
