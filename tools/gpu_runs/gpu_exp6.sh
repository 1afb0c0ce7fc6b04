#!/bin/bash
# development: ffmpeg drop-in + new tests, decoder A/B (generic vs planar) on real testsrc2
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ffmpeg_dropin.py tests/test_gpu_parity.py -m gpu -q -k "dropin or drop_in or version4 or resident or several" > gpurun_out/pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest6.log
tail -8 gpurun_out/pytest6.log
B="python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e"
for b in 24 96 192; do
  $B --batch $b > gpurun_out/b6_planar_b$b.json 2>/dev/null
  FFGPU_DEC_GENERIC=1 $B --batch $b > gpurun_out/b6_generic_b$b.json 2>/dev/null
done
FFGPU_DEC_GENERIC=1 FFGPU_LANE_STRIDE=2 $B > gpurun_out/b6_generic_s2.json 2>/dev/null
FFGPU_DEC_GENERIC=1 $B --synth > gpurun_out/b6_generic_synth.json 2>/dev/null
$B --synth > gpurun_out/b6_planar_synth.json 2>/dev/null
for f in gpurun_out/b6_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d["value"],1), {k:round(v,2) for k,v in (d.get("kernel_ms_per_step") or {}).items() if k in ("code","decode","init_state")})
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
