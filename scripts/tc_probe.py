"""Bring-up probe for the tcgen05 kernels: every case runs in its own subprocess with a timeout so that a hung kernel
(mbarrier deadlock) costs seconds, not the GPU lease.  Usage: python scripts/tc_probe.py [case ...]"""
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASE_SRC = r'''
import ctypes, sys, torch
sys.path.insert(0, %(root)r)
lib = ctypes.CDLL(%(root)r + "/onnxstream_b200/csrc/libonnxstream_b200.so")
vp, i64, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
lib.osb_gemm.argtypes = [vp, vp, vp, vp, vp, i64, i64, i64, i64, i64, i64, i64, ci, ci, ci, vp]
lib.osb_conv2d.argtypes = [vp, vp, vp, vp, vp, i64, i64, i64, i64, ci, ci, ci, ci, ci, i64, i64, ci, ci, vp]
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
kind = %(kind)r
torch.manual_seed(0)
if kind == "gemm":
    batch, M, N, K, bt = %(args)r
    a = torch.randn(batch, M, K, device="cuda").half()
    b = (torch.randn(batch, N, K, device="cuda") if bt else torch.randn(batch, K, N, device="cuda")).half()
    c = torch.full((batch, M, N), float("nan"), device="cuda", dtype=torch.half)
    rc = lib.osb_gemm(a.data_ptr(), b.data_ptr(), c.data_ptr(), None, None, batch, M, N, K, M*K, N*K, M*N, bt, 2, 2, st)
    torch.cuda.synchronize()
    ref = a.double() @ (b.double().transpose(1, 2) if bt else b.double())
    err = (c.double() - ref).abs()
    print("rc", rc, "max_err", float(err.max()), "ref_max", float(ref.abs().max()), "nan", int(torch.isnan(c).sum()))
    if float(err.max()) > 0.5 or torch.isnan(c).any():
        # where are the errors? print a coarse map of 32x32 blocks
        e = err[0]
        mb, nb = min(e.shape[0], 256) // 32, min(e.shape[1], 256) // 32
        for i in range(mb):
            print(" ".join("%%6.1f" %% float(torch.nan_to_num(e[i*32:(i+1)*32, j*32:(j+1)*32], nan=999).max()) for j in range(nb)))
else:
    H, W, Cin, Cout, k, pad = %(args)r
    import torch.nn.functional as F
    x = torch.randn(H, W, Cin, device="cuda").half()
    w = (torch.randn(Cout, k, k, Cin, device="cuda") / (k*k*Cin) ** 0.5).half()
    y = torch.full((H, W, Cout), float("nan"), device="cuda", dtype=torch.half)
    rc = lib.osb_conv2d(x.data_ptr(), w.data_ptr(), None, None, y.data_ptr(), H, W, Cin, Cout, k, k, 1, pad, pad, H, W, 2, 2, st)
    torch.cuda.synchronize()
    ref = F.conv2d(x.double().permute(2, 0, 1)[None], w.double().permute(0, 3, 1, 2), None, padding=pad)[0].permute(1, 2, 0)
    err = (y.double() - ref).abs()
    print("rc", rc, "max_err", float(err.max()), "ref_max", float(ref.abs().max()), "nan", int(torch.isnan(y).sum()))
'''

CASES = {
    "gemm_k_128": ("gemm", (1, 128, 128, 64, 1)),
    "gemm_k_256": ("gemm", (1, 256, 256, 256, 1)),
    "gemm_mn_128": ("gemm", (1, 128, 128, 64, 0)),
    "gemm_mn_256": ("gemm", (1, 256, 256, 256, 0)),
    "gemm_k_tail": ("gemm", (1, 300, 136, 72, 1)),
    "gemm_mn_tail": ("gemm", (1, 300, 136, 72, 0)),
    "gemm_mn_batch": ("gemm", (3, 200, 64, 40, 0)),
    "gemm_big": ("gemm", (1, 4096, 1280, 1280, 0)),
    "conv3": ("conv", (16, 16, 64, 128, 3, 1)),
    "conv1": ("conv", (16, 16, 64, 128, 1, 0)),
    "conv3_w8": ("conv", (8, 8, 128, 128, 3, 1)),
    "conv3_w24": ("conv", (24, 40, 32, 40, 3, 1)),
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for n in names:
        kind, args = CASES[n]
        src = CASE_SRC % dict(root=ROOT, kind=kind, args=args)
        try:
            r = subprocess.run([sys.executable, "-c", src], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=90)
            print(f"== {n}: exit {r.returncode}\n{r.stdout.strip()[-1500:]}", flush=True)
        except subprocess.TimeoutExpired as e:
            print(f"== {n}: TIMEOUT (hung kernel?)\n{(e.stdout or b'')[-500:]}", flush=True)
