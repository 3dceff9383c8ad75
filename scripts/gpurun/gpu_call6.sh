#!/bin/bash
mkdir -p gpurun_out
R1=$PWD/onnxstream_b200/csrc/libonnxstream_b200_r1.so
{
echo "--- r1 library"; OSB_ENGINE_LIB=$R1 timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
echo "--- default"; timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
echo "--- OSB_SIDE_BRANCH=0"; OSB_SIDE_BRANCH=0 timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
} > gpurun_out/r02_c6_ab.txt 2>&1
timeout 900 python -m pytest tests/test_models_gpu.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r02_c6_models.txt
OSB_SIDE_BRANCH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_step.csv python scripts/profile_step.py > gpurun_out/r02_c6_ncu.log 2>&1
OSB_ENGINE_LIB=$R1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_step_r1lib.csv python scripts/profile_step.py > gpurun_out/r02_c6_ncu_r1.log 2>&1
cat gpurun_out/r02_c6_ab.txt gpurun_out/r02_c6_models.txt; tail -3 gpurun_out/r02_c6_ncu.log gpurun_out/r02_c6_ncu_r1.log
