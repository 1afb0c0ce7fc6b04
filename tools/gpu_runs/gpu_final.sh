#!/bin/bash
# round 2: the record -- bench lines for every workload x source, reference arm, ncu launch
# list and full captures of the top kernels (summarised into profiles/ by profiles/summarize.py)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "two_pass or full_size or pipelined or dropin or staging or version4 or glue" > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final.log
tail -4 gpurun_out/pytest_final.log
python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/ref.err
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/n1.err; tail -2 gpurun_out/n1.err
for src in mandelbrot noise; do
  extra=""; [ $src = noise ] && extra="--batch 64"
  python bench.py --source $src $extra --steps 2 --warmup 3 --e2e-repeat 3 > gpurun_out/r02_bench_C2_$src.json 2> gpurun_out/$src.err; tail -2 gpurun_out/$src.err
done
for wl in C1 C3 C4 C5; do
  timeout 900 python bench.py --workload $wl --steps 2 --warmup 3 --e2e-repeat 1 --e2e-depth 3 > gpurun_out/r02_bench_$wl.json 2> gpurun_out/$wl.err; echo "$wl rc=$?"; tail -2 gpurun_out/$wl.err
done
for f in gpurun_out/r02_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    e=d.get("e2e") or {}
    print(sys.argv[1], "value", round(d["value"],1), "e2e", e.get("value"), "pg", (e.get("pageable") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"), {k:round(v,2) for k,v in (d.get("kernel_ms_per_step") or {}).items() if k in ("symbolize","code","decode")})
except Exception as ex:
    print(sys.argv[1], "ERR", ex)
PY
done
P="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e"
$P > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $P > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_decode|k_code_range|k_symbolize' -s 9 -c 3 -o gpurun_out/prof_r02 $P > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
P1="python bench.py --workload C1 --steps 1 --warmup 3 --no-cpu --no-e2e"
$P1 > gpurun_out/plain_c1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_decode|k_code_golomb|k_symbolize' -s 9 -c 3 -o gpurun_out/prof_r02_c1 $P1 > gpurun_out/ncu_full_c1.log 2>&1
echo "ncu c1 rc=$?"
