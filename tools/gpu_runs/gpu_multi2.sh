#!/bin/bash
# development: the routing handle on two real GPUs, and the 1 -> 2 curve both ways
set -x
mkdir -p gpurun_out
nvidia-smi -L; nvidia-smi topo -m | head -12
timeout 600 python -m pytest tests -m gpu -q -k "several or two_gpus" > gpurun_out/pytest_m2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_m2.log
tail -5 gpurun_out/pytest_m2.log
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/m2_n1.json 2> gpurun_out/m2_n1.err; tail -2 gpurun_out/m2_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/m2_n2.json 2> gpurun_out/m2_n2.err; tail -2 gpurun_out/m2_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-numa --no-pageable > gpurun_out/m2_n2_nonuma.json 2> gpurun_out/m2_n2_nonuma.err
python bench.py --steps 3 --warmup 3 --no-cpu --route 2 --no-pageable > gpurun_out/m2_route2.json 2> gpurun_out/m2_route2.err; tail -2 gpurun_out/m2_route2.err
for f in gpurun_out/m2_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    e=d.get("e2e") or {}
    print(sys.argv[1], "value", round(d["value"],1), "e2e", e.get("value"), "frac", e.get("fraction_of_all_ranks_duplex_probe"), "pg", (e.get("pageable") or {}).get("value"), "pcie", d.get("pcie"), d.get("numa"))
except Exception as ex:
    print(sys.argv[1], "ERR", ex)
PY
done
