#!/bin/bash
# development: full parity suite, lane-stride sweep on real testsrc2, decoder A/B, CPU baseline
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
B="python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e"
L=$PWD/ffmpeg_ffv2_b200/build
$B > gpurun_out/b4_s1.json 2> gpurun_out/b4_s1.err; tail -2 gpurun_out/b4_s1.err
FFGPU_LIB=$L/libffgpu_decprev.so $B > gpurun_out/b4_decprev.json 2>/dev/null
FFGPU_DEC_GENERIC=1 $B > gpurun_out/b4_generic.json 2>/dev/null
FFGPU_LANE_STRIDE=2 $B > gpurun_out/b4_s2.json 2>/dev/null
FFGPU_LANE_STRIDE=4 $B > gpurun_out/b4_s4.json 2>/dev/null
FFGPU_LANE_STRIDE=8 $B > gpurun_out/b4_s8.json 2>/dev/null
FFGPU_LANE_STRIDE=2 $B --batch 96 > gpurun_out/b4_s2_b96.json 2>/dev/null
$B --batch 96 > gpurun_out/b4_s1_b96.json 2>/dev/null
$B --source noise --batch 48 > gpurun_out/b4_noise.json 2> gpurun_out/b4_noise.err; tail -2 gpurun_out/b4_noise.err
python bench.py --steps 3 --warmup 3 > gpurun_out/b4_full.json 2> gpurun_out/b4_full.err; tail -2 gpurun_out/b4_full.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/b4_ref.json 2> gpurun_out/b4_ref.err
for f in gpurun_out/b4_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d["value"],1), "enc", d.get("encode_fps"), "dec", d.get("decode_fps"), {k:round(v,2) for k,v in (d.get("kernel_ms_per_step") or {}).items() if k in ("symbolize","code","decode")}, "e2e", (d.get("e2e") or {}).get("value"), "pg", ((d.get("e2e") or {}).get("pageable") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
