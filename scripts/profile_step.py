"""One eager SD1.5-UNet step between cudaProfilerStart/Stop, for ncu (`--profile-from-start off`).
Usage: ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py [workload]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from onnxstream_b200 import emit  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "sd15_unet_fp16"
W = bench.make_workload(workload)
d, meta = bench.ensure_model(W)
inputs = W.inputs(0)
m = bench.make_engine_model(d, W, "ram+nocache", resident=True, graph=False)
bench.step_api(m, inputs, W.out_name)
later = W.later_inputs(inputs)      # what a caller pushes on every later step (the KV cache of a decode step stays in HBM: b200_keep_inputs)
bench.step_api(m, later, W.out_name)
m.lib.model_b200_profiler(1)
bench.step_api(m, later, W.out_name)
m.lib.model_b200_profiler(0)
print("profiled one step:", {k: v for k, v in m.stats().items() if k in ("kernel_launches", "tc_launches", "last_gpu_ms")})
