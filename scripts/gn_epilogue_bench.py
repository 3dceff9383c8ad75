"""Micro-benchmark: conv with / without the GroupNorm-statistics epilogue (osb_conv2d_ex), and the stats / apply kernels alone.
Usage (GPU box): python scripts/gn_epilogue_bench.py   [OSB_GN_DEBUG=1|2|3 bisects the epilogue]"""
import ctypes, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.environ.get("OSB_ENGINE_LIB") or ROOT + "/onnxstream_b200/csrc/libonnxstream_b200.so")
vp, i64, ci, cf = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float
lib.osb_conv2d_ex.argtypes = [vp, vp, vp, vp, vp, vp, i64, i64, i64, i64, ci, ci, ci, ci, ci, i64, i64, ci, ci, vp, vp, ci, ctypes.POINTER(ci)]
lib.osb_group_norm_apply.argtypes = [vp, vp, ci, i64, i64, ci, vp, vp, cf, ci, vp, vp, vp]
lib.osb_channel_add_stats.argtypes = [vp, vp, vp, ci, i64, i64, ci, vp, vp]
lib.osb_group_norm.argtypes = [vp, vp, ci, ci, i64, i64, ci, vp, vp, cf, ci, vp, vp]
S = torch.cuda.Stream()
st = ctypes.c_void_p(S.cuda_stream)      # every launch goes to this stream: eager warm-up, then captured into a CUDA graph of `iters` launches

def timeit(fn, iters=50):
    with torch.cuda.stream(S):
        for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=S):
        for _ in range(iters): fn()
    g.replay(); torch.cuda.synchronize()
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1000.0

for (H, W, Cin, Cout, k) in [(64, 64, 320, 320, 3), (32, 32, 640, 640, 3), (16, 16, 1280, 1280, 3), (64, 64, 320, 320, 1)]:
    x = torch.randn(H, W, Cin, device="cuda").half(); w = (torch.randn(Cout, k, k, Cin, device="cuda") / (k*k*Cin) ** 0.5).half()
    b = torch.randn(Cout, device="cuda").half(); y = torch.empty(H, W, Cout, device="cuda", dtype=torch.half); o = torch.empty_like(y)
    stats = torch.zeros(64, device="cuda", dtype=torch.float64); stats2 = torch.zeros(64, device="cuda", dtype=torch.float64); scratch = torch.zeros(2048, device="cuda", dtype=torch.uint8)
    gamma = torch.ones(Cout, device="cuda").half(); beta = torch.zeros(Cout, device="cuda").half()
    done = ci(0)
    conv = lambda s: lib.osb_conv2d_ex(x.data_ptr(), w.data_ptr(), b.data_ptr(), None, None, y.data_ptr(), H, W, Cin, Cout, k, k, 1, k // 2, k // 2, H, W, 2, 0, st, s, 32 if s else 0, ctypes.byref(done))
    t0 = timeit(lambda: conv(None)); t1 = timeit(lambda: conv(stats.data_ptr()))
    ta = timeit(lambda: lib.osb_group_norm_apply(y.data_ptr(), o.data_ptr(), 2, Cout, H * W, 32, gamma.data_ptr(), beta.data_ptr(), 1e-5, 1, stats.data_ptr(), stats2.data_ptr(), st))
    ts = timeit(lambda: lib.osb_channel_add_stats(y.data_ptr(), None, None, 2, Cout, H * W, 32, stats.data_ptr(), st))
    tadd = timeit(lambda: lib.osb_channel_add_stats(y.data_ptr(), b.data_ptr(), o.data_ptr(), 2, Cout, H * W, 32, stats.data_ptr(), st))
    tf = timeit(lambda: lib.osb_group_norm(y.data_ptr(), o.data_ptr(), 2, 1, Cout, H * W, 32, gamma.data_ptr(), beta.data_ptr(), 1e-5, 1, scratch.data_ptr(), st))
    talt = timeit(lambda: (conv(None), conv(stats.data_ptr()))) / 2
    print(f"GNBENCH {H}x{W} {Cin}->{Cout} k{k}: conv {t0:.1f} us, conv+stats {t1:.1f} us (done={done.value}), apply {ta:.1f}, stats-only {ts:.1f}, add+stats {tadd:.1f}, r1 fused GN {tf:.1f}, alternating lean/stats conv {talt:.1f} per launch  carveout={os.environ.get('OSB_SMEM_CARVEOUT')} dbg={os.environ.get('OSB_GN_DEBUG')}", flush=True)
