"""tcgen05 kernel micro-benchmarks (large convs / GEMMs of the UNet plus 4096^3 and 8192^3): us, TF/s and weight GB/s per case."""
import ctypes, sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(ROOT + "/onnxstream_b200/csrc/libonnxstream_b200.so")
vp, i64, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
lib.osb_gemm.argtypes = [vp, vp, vp, vp, vp, i64, i64, i64, i64, i64, i64, i64, ci, ci, ci, vp]
lib.osb_conv2d.argtypes = [vp, vp, vp, vp, vp, i64, i64, i64, i64, ci, ci, ci, ci, ci, i64, i64, ci, ci, vp]
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1000.0

def conv(H, W, Cin, Cout, k):
    x = torch.randn(H, W, Cin, device="cuda").half()
    w = (torch.randn(Cout, k, k, Cin, device="cuda") / (k*k*Cin) ** 0.5).half()
    y = torch.empty(H, W, Cout, device="cuda", dtype=torch.half)
    us = timeit(lambda: lib.osb_conv2d(x.data_ptr(), w.data_ptr(), None, None, y.data_ptr(), H, W, Cin, Cout, k, k, 1, k // 2, k // 2, H, W, 2, 2, st))
    fl = 2.0 * H * W * Cin * Cout * k * k
    print(f"conv {H}x{W} {Cin}->{Cout} k{k}: {us:.1f} us  {fl / us * 1e-6:.1f} TF/s  w {Cout*k*k*Cin*2/us*1e-3:.0f} GB/s", flush=True)

def gemm(M, N, K, bt=1):
    a = torch.randn(1, M, K, device="cuda").half()
    b = (torch.randn(1, N, K, device="cuda") if bt else torch.randn(1, K, N, device="cuda")).half()
    c = torch.empty(1, M, N, device="cuda", dtype=torch.half)
    us = timeit(lambda: lib.osb_gemm(a.data_ptr(), b.data_ptr(), c.data_ptr(), None, None, 1, M, N, K, M*K, N*K, M*N, bt, 2, 2, st))
    fl = 2.0 * M * N * K
    print(f"gemm M{M} N{N} K{K} bt{bt}: {us:.1f} us  {fl / us * 1e-6:.1f} TF/s", flush=True)

print("env BN", os.environ.get("OSB_TC_BN"), "SPLIT", os.environ.get("OSB_TC_SPLIT"))
conv(32, 32, 1920, 640, 3)
gemm(1024, 640, 17280)
conv(64, 64, 640, 640, 3)
gemm(4096, 640, 5760)
conv(16, 16, 2560, 1280, 3)
conv(8, 8, 2560, 1280, 3)
gemm(4096, 2560, 320)
gemm(4096, 320, 1280)
gemm(4096, 4096, 4096)
gemm(8192, 8192, 8192)
gemm(8192, 8192, 8192, 0)
conv(128, 128, 320, 320, 3)      # SDXL 1024x1024 first level
conv(128, 128, 640, 320, 3)
conv(256, 256, 256, 256, 3)      # VAE decoder
conv(512, 512, 128, 128, 3)
conv(64, 64, 320, 320, 3)
conv(32, 32, 640, 640, 3)
