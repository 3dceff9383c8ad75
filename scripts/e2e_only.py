"""Quick A/B harness: the streamed e2e arm of bench.py only (host buffers, all weights through the HBM ring).
Usage: [OSB_* env toggles] python scripts/e2e_only.py [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
W = bench.make_workload("sd15_unet_fp16")
d, meta = bench.ensure_model(W)
inputs = W.inputs(0)
m = bench.make_engine_model(d, W, "ram+nocache", resident=False, graph=False)
for _ in range(2):
    bench.step_api(m, inputs, W.out_name)
t = time.perf_counter()
for _ in range(steps):
    bench.step_api(m, inputs, W.out_name)
dt = (time.perf_counter() - t) / steps
st = m.stats()
print(f"E2E_ONLY ms_per_step={dt * 1e3:.2f} h2d_gbs={st['weight_bytes_streamed'] / dt / 1e9:.1f} gpu_ms={st['last_gpu_ms']:.2f} run_ms={st['last_run_ms']:.2f} env={ {k: v for k, v in os.environ.items() if k.startswith('OSB_')} }", flush=True)
