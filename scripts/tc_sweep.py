"""Tuning sweep: time every unique tcgen05 GEMM / conv shape of the SD 1.5 UNet step (scripts/tc_shapes_sd15.txt: M N K taps batch conv count)
under the current OSB_TC_SPLIT / OSB_TC_BN overrides, cold L2 between launches.  One process per setting (the overrides are read once):
  for s in auto 1 2 3 4 6 8; do OSB_TC_SPLIT=$s python scripts/tc_sweep.py; done
The round-1 split-K / tile-width heuristics in gemm_tcgen05.cu come from this sweep."""
import ctypes, sys, os, torch, collections, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(ROOT + "/onnxstream_b200/csrc/libonnxstream_b200.so")
vp, i64, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
lib.osb_gemm.argtypes = [vp, vp, vp, vp, vp, i64, i64, i64, i64, i64, i64, i64, ci, ci, ci, vp]
lib.osb_conv2d.argtypes = [vp, vp, vp, vp, vp, i64, i64, i64, i64, ci, ci, ci, ci, ci, i64, i64, ci, ci, vp]
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
rows = [l.split() for l in open(ROOT + "/scripts/tc_shapes_sd15.txt") if l.strip()]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timeit(fn, iters=8):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()                       # cold L2, like a step
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters * 1000.0

out = []
for r in rows:
    M, N, K, taps, batch, conv, cnt = map(int, r)
    if conv:
        H = int(round(math.sqrt(M))); k = 3 if taps == 9 else 1
        x = torch.randn(H, H, K, device="cuda").half(); w = (torch.randn(N, k, k, K, device="cuda") / (k*k*K) ** 0.5).half()
        y = torch.empty(H, H, N, device="cuda", dtype=torch.half)
        us = timeit(lambda: lib.osb_conv2d(x.data_ptr(), w.data_ptr(), None, None, y.data_ptr(), H, H, K, N, k, k, 1, k // 2, k // 2, H, H, 2, 2, st))
    else:
        a = torch.randn(batch, M, K, device="cuda").half(); b = torch.randn(batch, N, K, device="cuda").half()
        c = torch.empty(batch, M, N, device="cuda", dtype=torch.half)
        us = timeit(lambda: lib.osb_gemm(a.data_ptr(), b.data_ptr(), c.data_ptr(), None, None, batch, M, N, K, M*K, N*K, M*N, 1, 2, 2, st))
    out.append(us)
print("RES", os.environ.get("OSB_TC_SPLIT", "auto"), " ".join(f"{u:.1f}" for u in out), flush=True)
