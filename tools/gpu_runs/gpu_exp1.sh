#!/bin/bash
# development: first GPU measurement of the specialised planar decoder (A/B against the generic one)
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
B="python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e"
$B > gpurun_out/b_new.json 2> gpurun_out/b_new.err; tail -2 gpurun_out/b_new.err
FFGPU_DEC_GENERIC=1 $B > gpurun_out/b_generic.json 2>/dev/null
FFGPU_GATE_WAIT=8 $B > gpurun_out/b_w8.json 2>/dev/null
FFGPU_GATE_WAIT=48 $B > gpurun_out/b_w48.json 2>/dev/null
FFGPU_GATE_WAIT=0 $B > gpurun_out/b_w0.json 2>/dev/null
FFGPU_LIB=$PWD/ffmpeg_ffv2_b200/build/libffgpu_ahead2.so $B > gpurun_out/b_ahead2.json 2>/dev/null
$B --batch 24 > gpurun_out/b_new_b24.json 2>/dev/null
FFGPU_LANE_STRIDE=1 $B --batch 24 > gpurun_out/b_new_b24_s1.json 2>/dev/null
for f in gpurun_out/b_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d["value"]), d["kernel_ms_per_step"], d.get("decisions"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
# source-level profile of the decode kernel (one launch)
$B --steps 1 --warmup 1 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_decode -s 2 -c 1 -o gpurun_out/prof_dec $B --steps 1 --warmup 1 > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"
