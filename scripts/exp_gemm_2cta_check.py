"""Bring-up harness for the EXPERIMENTAL cta_group::2 GEMM (onnxstream_b200/csrc/experimental/gemm_2cta.cu; not part of the product).
Builds it into build/libexp_gemm_2cta.so, then runs each size in its own subprocess with a timeout (a protocol bug in a pair kernel
shows up as a hang or a watchdog trap): correctness vs fp64 on the fp16-rounded operands, then time vs the product's 1-CTA kernel.
Usage (GPU box):  timeout 300 python scripts/exp_gemm_2cta_check.py [MxNxK ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "build", "libexp_gemm_2cta.so")
SRC = os.path.join(ROOT, "onnxstream_b200", "csrc", "experimental", "gemm_2cta.cu")

CASE = r'''
import ctypes, sys, torch
M, N, K = %(M)d, %(N)d, %(K)d
exp = ctypes.CDLL(%(so)r)
prod = ctypes.CDLL(%(root)r + "/onnxstream_b200/csrc/libonnxstream_b200.so")
vp, i64, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
exp.osb_exp_gemm_2cta.argtypes = [vp, vp, vp, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, vp]
prod.osb_gemm.argtypes = [vp, vp, vp, vp, vp, i64, i64, i64, i64, i64, i64, i64, ci, ci, ci, vp]
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
torch.manual_seed(0)
a = torch.randn(M, K, device="cuda").half(); b = torch.randn(N, K, device="cuda").half()
c = torch.full((M, N), float("nan"), device="cuda", dtype=torch.half)
rc = exp.osb_exp_gemm_2cta(a.data_ptr(), b.data_ptr(), c.data_ptr(), M, N, K, st)
torch.cuda.synchronize()
ref = (a[:512].double() @ b.double().t())
err = (c[:512].double() - ref).abs().max().item()
tol = (a[:512].double().abs() @ b.double().abs().t()).max().item() * 2 ** -9
print("rc", rc, "max_err(first 512 rows)", err, "tol", tol, "nan", int(torch.isnan(c).sum()), "OK" if rc == 0 and err <= tol and not torch.isnan(c).any() else "WRONG")
def timeit(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / it * 1e3
t2 = timeit(lambda: exp.osb_exp_gemm_2cta(a.data_ptr(), b.data_ptr(), c.data_ptr(), M, N, K, st))
t1 = timeit(lambda: prod.osb_gemm(a.data_ptr(), b.data_ptr(), c.data_ptr(), None, None, 1, M, N, K, M*K, N*K, M*N, 1, 2, 2, st))
fl = 2.0 * M * N * K
print(f"2-CTA {t2:.1f} us = {fl / t2 * 1e-6:.0f} TF/s   1-CTA {t1:.1f} us = {fl / t1 * 1e-6:.0f} TF/s")
'''


def main():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-o", SO, SRC, "-cudart", "static"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        print(r.stdout)
        sys.exit(1)
    sizes = sys.argv[1:] or ["256x256x64", "256x256x512", "512x768x1024", "4096x4096x4096", "8192x8192x8192"]
    for s in sizes:
        M, N, K = (int(v) for v in s.split("x"))
        src = CASE % dict(M=M, N=N, K=K, so=SO, root=ROOT)
        try:
            r = subprocess.run([sys.executable, "-c", src], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=60)
            print(f"== {s}: exit {r.returncode}\n{r.stdout.strip()[-1200:]}", flush=True)
        except subprocess.TimeoutExpired as e:
            print(f"== {s}: TIMEOUT (hung pair kernel?)\n{(e.stdout or b'')[-400:]}", flush=True)
            break


if __name__ == "__main__":
    main()
