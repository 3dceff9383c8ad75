#!/bin/bash
# round-2 GPU call 2: new full-size parity tests, whole GPU suite, default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_fullsize_gpu.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r02_c2_fullsize.txt
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_fullsize_gpu.py 2>&1 | tail -15 > gpurun_out/r02_c2_suite.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_c2_bench.json 2> gpurun_out/r02_c2_bench.err
tail -5 gpurun_out/r02_c2_bench.err
cat gpurun_out/r02_c2_fullsize.txt gpurun_out/r02_c2_suite.txt; cat gpurun_out/r02_c2_bench.json
