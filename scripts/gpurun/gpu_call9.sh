#!/bin/bash
mkdir -p gpurun_out
{
for d in 0 3; do OSB_GN_DEBUG=$d timeout 120 python scripts/gn_epilogue_bench.py 2>&1 | grep -E "GNBENCH|rror"; done
OSB_SMEM_CARVEOUT=1 timeout 120 python scripts/gn_epilogue_bench.py 2>&1 | grep -E "GNBENCH|rror"
echo "--- value default"; timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
echo "--- value OSB_SMEM_CARVEOUT=1"; OSB_SMEM_CARVEOUT=1 timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
echo "--- value OSB_GN_SPLIT=0"; OSB_GN_SPLIT=0 timeout 200 python scripts/value_only.py 20 2>&1 | grep -E "VALUE_ONLY|rror"
} > gpurun_out/r02_c9_gnbench.txt 2>&1
cat gpurun_out/r02_c9_gnbench.txt
