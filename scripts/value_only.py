"""Quick A/B harness: the resident-weights + CUDA-graph arm of bench.py only (no e2e, no CPU legs).
Usage: [OSB_* env toggles] python scripts/value_only.py [steps]   -> prints ms per UNet step (CUDA-event time of the replays)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from onnxstream_b200 import emit  # noqa: E402

# note: importing bench redirects fd 1 to stderr (its JSON contract); this script only writes diagnostics, so that is fine
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
W = bench.make_workload("sd15_unet_fp16")
d, meta = bench.ensure_model(W)
inputs = W.inputs(0)
m = bench.make_engine_model(d, W, "ram+nocache", True, True)
for _ in range(4):
    bench.step_api(m, inputs, W.out_name)
ms = m.run_resident(steps) / steps
st = m.stats()
print(f"VALUE_ONLY ms_per_step={ms:.4f} launches={st.get('kernel_launches')} tc={st.get('tc_launches')} side_steps={st.get('side_steps')} env={ {k: v for k, v in os.environ.items() if k.startswith('OSB_')} }", flush=True)
if os.environ.get("OSB_TC_DUMP"):
    import ctypes
    lib = m.lib
    m.lib.model_set_option(m.h, b"b200_cuda_graph", 0)
    lib.osb_tc_profile(1)
    bench.step_api(m, inputs, W.out_name)
    buf = ctypes.create_string_buffer(1 << 20)
    lib.osb_tc_profile_dump.argtypes = [ctypes.c_char_p, ctypes.c_int]
    n = lib.osb_tc_profile_dump(buf, len(buf))
    open(os.environ["OSB_TC_DUMP"], "wb").write(buf.raw[:max(n, 0)])
    lib.osb_tc_profile(0)
