#!/bin/bash
# 4 GPUs: bench --gpus 4 (sharded upload + all-gather over 4 ranks, all-rank parity)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02_c17_bench_n4.json 2> gpurun_out/r02_c17_bench_n4.err
tail -n 3 gpurun_out/r02_c17_bench_n4.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r02_c17_bench_n4.json")); print("N4", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("h2d_gbs"), d["parity"])
except Exception as e: print("ERR", e)
PY
