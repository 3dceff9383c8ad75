#!/bin/bash
# 2 GPUs: multi-GPU parity tests + bench lines (default workload and the SDXL CFG pair)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_c12_gpus.txt
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r02_c12_mgpu_tests.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_c12_bench_n2.json 2> gpurun_out/r02_c12_bench_n2.err
OSB_SHARDED_H2D=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c12_bench_n2_bcast.json 2> gpurun_out/r02_c12_bench_n2_bcast.err
cat gpurun_out/r02_c12_mgpu_tests.txt; tail -n 4 gpurun_out/r02_c12_bench_n2.err; cat gpurun_out/r02_c12_bench_n2.json | cut -c1-1500; cat gpurun_out/r02_c12_bench_n2_bcast.json | cut -c1-600
