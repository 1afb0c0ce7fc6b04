#!/bin/bash
# development: parity + kernel A/B after the state-pipelining change, first lines of the other workloads
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
B="python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e"
L=$PWD/ffmpeg_ffv2_b200/build
$B --synth > gpurun_out/b3_synth.json 2> gpurun_out/b3_synth.err; tail -2 gpurun_out/b3_synth.err
FFGPU_LIB=$L/libffgpu_mb6.so $B --synth > gpurun_out/b3_synth_mb6.json 2>/dev/null
$B > gpurun_out/b3_testsrc2.json 2> gpurun_out/b3_testsrc2.err; tail -2 gpurun_out/b3_testsrc2.err
$B --source mandelbrot > gpurun_out/b3_mandelbrot.json 2>/dev/null
$B --source noise --batch 96 > gpurun_out/b3_noise.json 2> gpurun_out/b3_noise.err; tail -2 gpurun_out/b3_noise.err
W="python bench.py --steps 2 --warmup 1 --no-e2e"
for wl in C1 C3 C4 C5; do
  timeout 600 $W --workload $wl > gpurun_out/b3_$wl.json 2> gpurun_out/b3_$wl.err; echo "$wl rc=$?"; tail -2 gpurun_out/b3_$wl.err
done
for f in gpurun_out/b3_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], round(d["value"],1), "enc", d.get("encode_fps"), "dec", d.get("decode_fps"), {k:round(v,2) for k,v in d["kernel_ms_per_step"].items() if k in ("symbolize","code","decode")}, "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
nvidia-smi --query-gpu=memory.used --format=csv
