#!/bin/bash
# round-2 GPU call 1: 2-CTA skeleton bring-up, PDL A/B, full-size reference timing
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_c1_env.txt; nproc >> gpurun_out/r02_c1_env.txt
timeout 300 python scripts/exp_gemm_2cta_check.py > gpurun_out/r02_2cta_check.txt 2>&1
echo "--- baseline" > gpurun_out/r02_pdl_ab.txt
timeout 200 python scripts/value_only.py 20 2>&1 | grep VALUE_ONLY >> gpurun_out/r02_pdl_ab.txt
echo "--- OSB_PDL=1 (trigger at entry)" >> gpurun_out/r02_pdl_ab.txt
OSB_PDL=1 timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror" >> gpurun_out/r02_pdl_ab.txt
echo "--- OSB_PDL=1 late trigger variant" >> gpurun_out/r02_pdl_ab.txt
OSB_PDL=1 OSB_ENGINE_LIB=$PWD/onnxstream_b200/csrc/libonnxstream_b200_pdllate.so timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror" >> gpurun_out/r02_pdl_ab.txt
echo "--- late variant without PDL attr" >> gpurun_out/r02_pdl_ab.txt
OSB_ENGINE_LIB=$PWD/onnxstream_b200/csrc/libonnxstream_b200_pdllate.so timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror" >> gpurun_out/r02_pdl_ab.txt
timeout 400 python scripts/ref_full_time.py 32 2>&1 | grep -E "REF_FULL|PARITY|rror" > gpurun_out/r02_ref_full.txt
cat gpurun_out/r02_2cta_check.txt gpurun_out/r02_pdl_ab.txt gpurun_out/r02_ref_full.txt
