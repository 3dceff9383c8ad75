#!/bin/bash
# 1 GPU: all kernel tests (new: grouped GEMV, f32x, concat2), model tests, decode / CLIP bench lines, i8 + f32x ncu captures, llama launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu 2>&1 | tail -12 > gpurun_out/r02_c14_ktests.txt
timeout 1200 python -m pytest tests/test_models_gpu.py -x -q -m gpu 2>&1 | tail -12 > gpurun_out/r02_c14_tests.txt
cat gpurun_out/r02_c14_ktests.txt gpurun_out/r02_c14_tests.txt
for w in tinyllama_decode tinyllama_decode_w8 clip_text_fp32; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c14_$w.json 2> gpurun_out/r02_c14_$w.err
  tail -n 2 gpurun_out/r02_c14_$w.err
done
python - <<'PY'
import json
for w in ["tinyllama_decode","tinyllama_decode_w8","clip_text_fp32"]:
    try:
        d=json.load(open(f"gpurun_out/r02_c14_{w}.json")); print(w, d["value"], d["unit"], d["ms_per_step"], d.get("gpu_launches_per_step"), d["roofline"]["frac"], d["e2e"]["value"], d.get("parity"))
    except Exception as e: print(w, "ERR", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_llama_after.csv python scripts/profile_step.py tinyllama_decode > gpurun_out/r02_c14_ncu2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_clip.csv python scripts/profile_step.py clip_text_fp32 > gpurun_out/r02_c14_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_i8_kernel" -c 3 -o gpurun_out/r02_full_i8 -f python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "qu8_tensor_core" > gpurun_out/r02_c14_ncu5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"attention_decode_kernel|gemv|rms_norm" --launch-skip 6 -c 8 -o gpurun_out/r02_full_llama_after -f python scripts/profile_step.py tinyllama_decode > gpurun_out/r02_c14_ncu6.log 2>&1
python scripts/summarize_launches.py gpurun_out/r02_launches_llama_after.csv 2>&1 | head -30
python scripts/summarize_launches.py gpurun_out/r02_launches_clip.csv 2>&1 | head -30
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_c14_bench_default.json 2> gpurun_out/r02_c14_bench_default.err
tail -n 3 gpurun_out/r02_c14_bench_default.err; cut -c1-700 gpurun_out/r02_c14_bench_default.json
