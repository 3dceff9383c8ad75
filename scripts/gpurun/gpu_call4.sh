#!/bin/bash
# round-2 GPU call 4 (1 GPU): single-kernel epilogue A/B, model + full-size suites with the pair kernel / GN statistics / qdq, bench line
mkdir -p gpurun_out
{
echo "--- noepi lib, PAIR=0 GN_SPLIT=0"; OSB_ENGINE_LIB=$PWD/onnxstream_b200/csrc/libonnxstream_b200_noepi.so OSB_TC_PAIR=0 OSB_GN_SPLIT=0 timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
echo "--- default lib, PAIR=0 GN_SPLIT=0"; OSB_TC_PAIR=0 OSB_GN_SPLIT=0 timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
echo "--- default lib"; timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
echo "--- noepi lib again"; OSB_ENGINE_LIB=$PWD/onnxstream_b200/csrc/libonnxstream_b200_noepi.so OSB_TC_PAIR=0 OSB_GN_SPLIT=0 timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
} > gpurun_out/r02_c4_ab.txt 2>&1
timeout 900 python -m pytest tests/test_models_gpu.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r02_c4_models.txt
timeout 1500 python -m pytest tests/test_fullsize_gpu.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r02_c4_fullsize.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_c4_bench.json 2> gpurun_out/r02_c4_bench.err
cat gpurun_out/r02_c4_ab.txt gpurun_out/r02_c4_models.txt gpurun_out/r02_c4_fullsize.txt; tail -3 gpurun_out/r02_c4_bench.err; cat gpurun_out/r02_c4_bench.json
