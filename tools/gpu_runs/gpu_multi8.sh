#!/bin/bash
# round 2: the 1 -> 4 -> 8 curve (torchrun ranks) and the product's own routing handle over 8 GPUs
set -x
mkdir -p gpurun_out
nvidia-smi topo -m | head -14
python bench.py --steps 3 --warmup 3 --no-cpu --no-pageable > gpurun_out/m8_n1.json 2> gpurun_out/m8_n1.err; tail -2 gpurun_out/m8_n1.err
for n in 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 --no-pageable > gpurun_out/m8_n$n.json 2> gpurun_out/m8_n$n.err; tail -2 gpurun_out/m8_n$n.err
done
python bench.py --steps 3 --warmup 3 --no-cpu --route 8 --no-pageable > gpurun_out/m8_route8.json 2> gpurun_out/m8_route8.err; tail -2 gpurun_out/m8_route8.err
timeout 300 python -m pytest tests -m gpu -q -k "several" > gpurun_out/pytest_m8.log 2>&1; tail -2 gpurun_out/pytest_m8.log
for f in gpurun_out/m8_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    e=d.get("e2e") or {}
    print(sys.argv[1], "value", round(d["value"],1), "e2e", e.get("value"), "frac", e.get("fraction_of_all_ranks_duplex_probe"), "agg", (d.get("pcie") or {}).get("all_ranks_duplex_GBps_each_way_aggregate"), d.get("numa"))
except Exception as ex:
    print(sys.argv[1], "ERR", ex)
PY
done
