#!/bin/bash
mkdir -p gpurun_out
timeout 100 python bench.py --workload clip_text_fp32 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c18_clip.json 2> gpurun_out/r02_c18_clip.err
python -c "import json; d=json.load(open('gpurun_out/r02_c18_clip.json')); print('clip', d['value'], d['roofline']['achieved'], d['roofline']['frac'])"
timeout 170 python bench.py --workload sd15_unet_fp32 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c18_fp32.json 2> gpurun_out/r02_c18_fp32.err
python -c "import json; d=json.load(open('gpurun_out/r02_c18_fp32.json')); print('fp32', d['value'], d['roofline']['achieved'], d['roofline']['frac'])"
