#!/bin/bash
mkdir -p gpurun_out
R1=$PWD/onnxstream_b200/csrc/libonnxstream_b200_r1.so
{
echo "--- r1 library"; OSB_ENGINE_LIB=$R1 timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
echo "--- default"; timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
echo "--- OSB_SIDE_BRANCH=0"; OSB_SIDE_BRANCH=0 timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
echo "--- OSB_TC_PAIR=0"; OSB_TC_PAIR=0 timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
echo "--- default again"; timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
} > gpurun_out/r02_c7_ab.txt 2>&1
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "gn or channel or pair" 2>&1 | tail -5 > gpurun_out/r02_c7_kernels.txt
OSB_SIDE_BRANCH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_step.csv python scripts/profile_step.py > gpurun_out/r02_c7_ncu.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_c7_bench.json 2> gpurun_out/r02_c7_bench.err
cat gpurun_out/r02_c7_ab.txt gpurun_out/r02_c7_kernels.txt; tail -n 3 gpurun_out/r02_c7_ncu.log; tail -n 3 gpurun_out/r02_c7_bench.err; cat gpurun_out/r02_c7_bench.json
