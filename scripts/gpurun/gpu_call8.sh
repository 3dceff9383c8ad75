#!/bin/bash
mkdir -p gpurun_out
R1=$PWD/onnxstream_b200/csrc/libonnxstream_b200_r1.so
{
echo "--- r1 library"; OSB_ENGINE_LIB=$R1 timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
echo "--- default"; timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
echo "--- OSB_GN_SPLIT=0"; OSB_GN_SPLIT=0 timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
echo "--- OSB_GN_SPLIT=0 OSB_TC_PAIR=0"; OSB_GN_SPLIT=0 OSB_TC_PAIR=0 timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
echo "--- OSB_SIDE_BRANCH=1"; OSB_SIDE_BRANCH=1 timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
echo "--- OSB_PDL=1"; OSB_PDL=1 timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
} > gpurun_out/r02_c8_ab.txt 2>&1
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r02_c8_tests.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_step_warm.csv python scripts/profile_step.py > gpurun_out/r02_c8_ncu.log 2>&1
OSB_ENGINE_LIB=$R1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_step_warm_r1lib.csv python scripts/profile_step.py > gpurun_out/r02_c8_ncu_r1.log 2>&1
cat gpurun_out/r02_c8_ab.txt gpurun_out/r02_c8_tests.txt; tail -n 2 gpurun_out/r02_c8_ncu.log
