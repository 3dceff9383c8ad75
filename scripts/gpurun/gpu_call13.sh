#!/bin/bash
# 1 GPU: new kernel tests, LLM / CLIP model tests, decode + CLIP bench lines with graph capture, ncu launch lists + full captures
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "attention_decode or rms_norm or rope" 2>&1 | tail -8 > gpurun_out/r02_c13_ktests.txt
timeout 900 python -m pytest tests/test_models_gpu.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r02_c13_tests.txt
for w in tinyllama_decode tinyllama_decode_w8 clip_text_fp32; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c13_$w.json 2> gpurun_out/r02_c13_$w.err
  tail -n 2 gpurun_out/r02_c13_$w.err
done
cat gpurun_out/r02_c13_ktests.txt gpurun_out/r02_c13_tests.txt
python - <<'PY'
import json
for w in ["tinyllama_decode","tinyllama_decode_w8","clip_text_fp32"]:
    try:
        d=json.load(open(f"gpurun_out/r02_c13_{w}.json")); print(w, d["value"], d["unit"], d["ms_per_step"], d.get("gpu_launches_per_step"), d["config"]["weights"][:90], d["roofline"]["frac"], d["e2e"]["value"], d.get("parity"))
    except Exception as e: print(w, "ERR", e)
PY
# launch lists (time + DRAM bytes per launch) of one eager step: SD1.5 UNet and the llama decode step
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_step.csv python scripts/profile_step.py > gpurun_out/r02_c13_ncu1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_llama.csv python scripts/profile_step.py tinyllama_decode > gpurun_out/r02_c13_ncu2.log 2>&1
# full captures: the CTA-pair kernel, the single-CTA kernel (split-K weight-streaming conv included), GN apply; then the i8 kernel from its test
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"tc_pair_kernel" -c 6 -o gpurun_out/r02_full_pair -f python scripts/profile_step.py > gpurun_out/r02_c13_ncu3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"tc_gemm_kernel|gn_apply_pre_kernel|splitk_reduce_kernel" --launch-skip 30 -c 10 -o gpurun_out/r02_full_gemm -f python scripts/profile_step.py > gpurun_out/r02_c13_ncu4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_i8_kernel" -c 3 -o gpurun_out/r02_full_i8 -f python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "qu8_tc" > gpurun_out/r02_c13_ncu5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"attention_decode_kernel|gemv" --launch-skip 8 -c 6 -o gpurun_out/r02_full_llama -f python scripts/profile_step.py tinyllama_decode > gpurun_out/r02_c13_ncu6.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_pair_kernel" -c 3 -o gpurun_out/r02_full_pair_test -f python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "pair" > gpurun_out/r02_c13_ncu7.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -8
python scripts/summarize_launches.py gpurun_out/r02_launches_step.csv 2>&1 | head -30
python scripts/summarize_launches.py gpurun_out/r02_launches_llama.csv 2>&1 | head -30
