#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_models_gpu.py -x -q -m gpu -k "llama or clip or exported" 2>&1 | tail -8 > gpurun_out/r02_c13_tests.txt
for w in tinyllama_decode tinyllama_decode_w8 clip_text_fp32; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c13_$w.json 2> gpurun_out/r02_c13_$w.err
  tail -n 2 gpurun_out/r02_c13_$w.err
done
cat gpurun_out/r02_c13_tests.txt
python - <<'PY'
import json
for w in ["tinyllama_decode","tinyllama_decode_w8","clip_text_fp32"]:
    try:
        d=json.load(open(f"gpurun_out/r02_c13_{w}.json")); print(w, d["value"], d["unit"], d["ms_per_step"], d["gpu_launches_per_step"], d["config"]["weights"][:90], d["roofline"]["frac"], d["e2e"]["value"])
    except Exception as e: print(w, "ERR", e)
PY
