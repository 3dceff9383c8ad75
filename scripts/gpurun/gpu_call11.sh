#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r02_c11_kernels.txt
{
echo "--- default (in-kernel last-finisher split-K, min 16 k-blocks)"; timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
echo "--- OSB_TC_SPLIT_MINKB=8"; OSB_TC_SPLIT_MINKB=8 timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
echo "--- OSB_TC_SPLIT_MINKB=12"; OSB_TC_SPLIT_MINKB=12 timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
echo "--- OSB_TC_SPLIT_MINKB=32"; OSB_TC_SPLIT_MINKB=32 timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
echo "--- OSB_TC_INKERNEL_REDUCE=0"; OSB_TC_INKERNEL_REDUCE=0 timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
} > gpurun_out/r02_c11_ab.txt 2>&1
timeout 900 python -m pytest tests/test_models_gpu.py -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r02_c11_models.txt
cat gpurun_out/r02_c11_kernels.txt gpurun_out/r02_c11_ab.txt gpurun_out/r02_c11_models.txt
