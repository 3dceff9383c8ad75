#!/bin/bash
# 2 GPUs: multi-GPU parity tests, bench --gpus 2 with the all-rank reference check (sharded upload + all-gather), SDXL CFG pair
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r02_c15_mgpu_tests.txt
cat gpurun_out/r02_c15_mgpu_tests.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_c15_bench_n2.json 2> gpurun_out/r02_c15_bench_n2.err
tail -n 3 gpurun_out/r02_c15_bench_n2.err | cut -c1-300; cut -c1-400 gpurun_out/r02_c15_bench_n2.json
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r02_c15_bench_n2.json")); print("N2", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["parity"])
except Exception as e: print("ERR", e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r02_c15_ref_n2.json 2> gpurun_out/r02_c15_ref_n2.err
cut -c1-300 gpurun_out/r02_c15_ref_n2.json
