#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 > gpurun_out/r02_c10_tests.txt
rm -f gpurun_out/r02_configs.json
timeout 2400 python bench.py --steps 10 --warmup 3 --all-configs gpurun_out/r02_configs.json > gpurun_out/r02_c10_bench_default.json 2> gpurun_out/r02_c10_bench.err
cat gpurun_out/r02_c10_tests.txt; tail -n 5 gpurun_out/r02_c10_bench.err; wc -l gpurun_out/r02_configs.json
