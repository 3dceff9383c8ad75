#!/bin/bash
# round-2 GPU call 5 (1 GPU): same-box A/B against the round-1 library, side-branch A/B, new kernel tests (kind::i8, gemv_w8), models, ncu launch list
mkdir -p gpurun_out
R1=$PWD/onnxstream_b200/csrc/libonnxstream_b200_r1.so
{
echo "--- r1 library (round-1 kernels + engine)"; OSB_ENGINE_LIB=$R1 timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
echo "--- default"; timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
echo "--- OSB_SIDE_BRANCH=0"; OSB_SIDE_BRANCH=0 timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
echo "--- OSB_SIDE_BRANCH=0 OSB_GN_SPLIT=0"; OSB_SIDE_BRANCH=0 OSB_GN_SPLIT=0 timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
echo "--- OSB_SIDE_BRANCH=0 OSB_GN_SPLIT=0 OSB_TC_PAIR=0"; OSB_SIDE_BRANCH=0 OSB_GN_SPLIT=0 OSB_TC_PAIR=0 timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
echo "--- r1 library again"; OSB_ENGINE_LIB=$R1 timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
} > gpurun_out/r02_c5_ab.txt 2>&1
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r02_c5_kernels.txt
timeout 900 python -m pytest tests/test_models_gpu.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r02_c5_models.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_step.csv python scripts/profile_step.py > gpurun_out/r02_c5_ncu.log 2>&1
cat gpurun_out/r02_c5_ab.txt gpurun_out/r02_c5_kernels.txt gpurun_out/r02_c5_models.txt; tail -3 gpurun_out/r02_c5_ncu.log
