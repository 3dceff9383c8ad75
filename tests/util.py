"""Shared helpers for the parity tests: run the same model directory + inputs through two libraries exporting the
reference's C ABI (the B200 engine and the reference-compiled oracle) and compare every requested tensor."""
import numpy as np

from onnxstream_b200.model import Model


def run_model(lib, model_dir, inputs, options=(), extra_outputs=(), wp="nocache", parts=None, ranges=None, upcast=(), b200_options=(), runs=1, plain_abi=False):
    m = Model(lib, 4, wp, plain_abi=plain_abi)   # 4 pthreadpool workers for the CPU oracle (0 = every core: slow on many-core hosts); the GPU engine ignores it
    for o in options:
        m.set_option(o, True)
    for name, val in b200_options:
        m.lib.model_set_option(m.h, name.encode(), int(val))
    if parts is not None:
        m.set_attention_parts(parts)
    for k, (mn, mx) in (ranges or {}).items():
        m.set_range(k, mn, mx)
    for p in upcast:
        m.add_upcast_pattern(p)
    for e in extra_outputs:
        m.add_extra_output(e)
    m.read_file(model_dir.rstrip("/") + "/model.txt")
    out = None
    for _ in range(runs):
        m.clear_tensors()
        for k, v in inputs.items():
            m.add_tensor(k, v)
        m.run()
        out = {}
        for n in m.get_all_tensor_names():
            t = m.get_tensor(n)
            if t is None and not plain_abi:
                t = m.get_tensor_i64(n)
            out[n] = t
    return out, m


def report(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    d = np.abs(a - b)
    denom = max(np.abs(b).max(), 1e-12)
    return dict(max_abs=float(d.max()) if d.size else 0.0, rel_to_max=float(d.max() / denom) if d.size else 0.0,
                rms=float(np.sqrt((d ** 2).mean())) if d.size else 0.0, ref_rms=float(np.sqrt((b ** 2).mean())) if b.size else 0.0)
