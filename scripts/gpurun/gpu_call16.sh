#!/bin/bash
# 1 GPU: the whole GPU suite, smoke(), every BASELINE workload, GEMV grid A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu 2>&1 | tail -25 > gpurun_out/r02_c16_ktests.txt
timeout 1200 python -m pytest tests/test_models_gpu.py -q -m gpu 2>&1 | tail -25 > gpurun_out/r02_c16_tests.txt
timeout 1500 python -m pytest tests/test_fullsize_gpu.py -q -m gpu 2>&1 | tail -25 > gpurun_out/r02_c16_fullsize.txt
cut -c1-300 gpurun_out/r02_c16_ktests.txt | tail -12; cut -c1-300 gpurun_out/r02_c16_tests.txt | tail -12; cut -c1-300 gpurun_out/r02_c16_fullsize.txt | tail -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
rm -f gpurun_out/r02_configs.json
timeout 1500 python bench.py --all-configs gpurun_out/r02_configs.json --steps 20 --warmup 3 > gpurun_out/r02_c16_bench_default.json 2> gpurun_out/r02_c16_bench.err
tail -n 3 gpurun_out/r02_c16_bench.err | cut -c1-300
python - <<'PY'
import json
for l in open("gpurun_out/r02_configs.json"):
    d=json.loads(l)
    if "error" in d: print(d); continue
    print(d.get("workload_key"), "value", round(d["value"],2), d["unit"], "ms", round(d["ms_per_step"],3), "e2e", d.get("e2e",{}).get("value"), "frac", d.get("roofline",{}).get("frac"), "parity", (d.get("parity") or {}).get("vs_reference_rel_all_ranks"), (d.get("parity") or {}).get("ok"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
PY
for c in 296 1184; do
  OSB_GEMV_CTAS=$c timeout 600 python bench.py --workload tinyllama_decode --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('GEMV_CTAS=$c', d['value'], d['ms_per_step'])"
done
