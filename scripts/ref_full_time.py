"""Time one FULL run of the reference CPU path (oracle/_ref) on the SD1.5 UNet-shaped graph with T threads and compare with the engine.
Usage: python scripts/ref_full_time.py [threads]"""
import os
import sys
import time

T = int(sys.argv[1]) if len(sys.argv) > 1 else 32
os.environ["OMP_NUM_THREADS"] = str(T)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import bench  # noqa: E402
from onnxstream_b200 import emit  # noqa: E402
from onnxstream_b200.model import Model  # noqa: E402

W = bench.make_workload("sd15_unet_fp16")
d, meta = bench.ensure_model(W)
inputs = W.inputs(0)
m = Model(bench.ORACLE_LIB, T, "nocache")
m.set_option("fuse_ops_in_attention", True)
m.read_file(d + "model.txt")
for it in range(2):
    t = time.time()
    ref = bench.step_api(m, inputs, W.out_name)
    print(f"REF_FULL threads={T} run{it} {time.time() - t:.2f} s", flush=True)
e = bench.make_engine_model(d, W, "ram+nocache", True, False)
got = bench.step_api(e, inputs, W.out_name)
print("PARITY max|err|", float(np.abs(got - ref).max()), "max|ref|", float(np.abs(ref).max()), flush=True)
