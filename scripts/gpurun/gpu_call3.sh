#!/bin/bash
# round-2 GPU call 3: pair kernel + GN-stats tests, e2e A/B (NUMA policy / grouping), step time, pair microbench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r02_c3_kernels.txt
timeout 900 python -m pytest tests/test_models_gpu.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r02_c3_models.txt
{
echo "--- value (default)"; timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
echo "--- value OSB_TC_PAIR=0"; OSB_TC_PAIR=0 timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
echo "--- value OSB_GN_SPLIT=0"; OSB_GN_SPLIT=0 timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
echo "--- value OSB_TC_PAIR=0 OSB_GN_SPLIT=0"; OSB_TC_PAIR=0 OSB_GN_SPLIT=0 timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
echo "--- e2e default"; timeout 200 python scripts/e2e_only.py 5 2>&1 | grep -E "E2E_ONLY|rror"
echo "--- e2e OSB_NUMA_POLICY=0"; OSB_NUMA_POLICY=0 timeout 200 python scripts/e2e_only.py 5 2>&1 | grep -E "E2E_ONLY|rror"
echo "--- e2e OSB_NUMA_BIND=0"; OSB_NUMA_BIND=0 timeout 200 python scripts/e2e_only.py 5 2>&1 | grep -E "E2E_ONLY|rror"
echo "--- e2e OSB_WEIGHT_GROUP_KB=0"; OSB_WEIGHT_GROUP_KB=0 timeout 200 python scripts/e2e_only.py 5 2>&1 | grep -E "E2E_ONLY|rror"
echo "--- e2e OSB_WEIGHT_GROUP_KB=32768"; OSB_WEIGHT_GROUP_KB=32768 timeout 200 python scripts/e2e_only.py 5 2>&1 | grep -E "E2E_ONLY|rror"
nvidia-smi topo -m 2>&1 | head -12; cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c | head
} > gpurun_out/r02_c3_ab.txt 2>&1
{
echo "--- microbench pair=1"; timeout 300 python scripts/tc_microbench.py 2>&1 | tail -14
echo "--- microbench pair=0"; OSB_TC_PAIR=0 timeout 300 python scripts/tc_microbench.py 2>&1 | tail -14
} > gpurun_out/r02_c3_micro.txt 2>&1
cat gpurun_out/r02_c3_kernels.txt gpurun_out/r02_c3_models.txt gpurun_out/r02_c3_ab.txt gpurun_out/r02_c3_micro.txt
