"""CUDA-event timing of representative GEMM / conv shapes of the UNet (warm, back to back, rotating
buffers so the activations do not live in L2).  Usage: python scripts/gemm_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from storygen_b200 import ops, _lib
from storygen_b200.ops import BF16, Src, taps3x3

dev = "cuda"

def timeit(fn, nbuf, iters=20):
    for i in range(3): fn(i % nbuf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fn(i % nbuf)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us

def linear_case(m, n, k, res=False, geglu=False):
    nbuf = max(1, min(6, int(300e6 // (m * (k + n) * 2)) or 1))
    xs = [torch.randn(m, k, device=dev).to(BF16) for _ in range(nbuf)]
    w = (torch.randn(n, k, device=dev) * k ** -0.5).to(BF16)
    b = torch.zeros(n, device=dev)
    nout = n // 2 if geglu else n
    ys = [torch.empty(m, nout, device=dev, dtype=BF16) for _ in range(nbuf)]
    rs = [torch.randn(m, nout, device=dev).to(BF16) for _ in range(nbuf)] if res else None
    mode = _lib.SG_EPI_GEGLU if geglu else _lib.SG_EPI_LINEAR
    us = timeit(lambda i: ops.linear(xs[i], w, ys[i], bias=b, residual=rs[i] if res else None, mode=mode), nbuf)
    fl = 2.0 * m * n * k
    by = (m * k + m * nout * (2 if res else 1)) * 2 + n * k * 2
    print(f"linear M={m:6d} N={n:5d} K={k:5d} res={int(res)} geglu={int(geglu)}: {us:8.1f} us  {fl/us/1e6:7.0f} TF/s  {by/us/1e3:6.0f} GB/s", flush=True)

def conv_case(n, h, w, cin, cout):
    nbuf = max(1, min(6, int(300e6 // (n * h * w * (cin + cout) * 2)) or 1))
    xs = [torch.randn(n, h, w, cin, device=dev).to(BF16) for _ in range(nbuf)]
    wt = (torch.randn(cout, 9 * cin, device=dev) * (9 * cin) ** -0.5).to(BF16)
    b = torch.zeros(cout, device=dev)
    ys = [torch.empty(n, h, w, cout, device=dev, dtype=BF16) for _ in range(nbuf)]
    us = timeit(lambda i: ops.gemm([Src.nhwc(xs[i])], taps3x3(0, cin), (w, h, n), wt, ys[i], bias=b), nbuf)
    fl = 2.0 * n * h * w * cout * 9 * cin
    print(f"conv3x3 n={n} {h}x{w} {cin}->{cout}: {us:8.1f} us  {fl/us/1e6:7.0f} TF/s", flush=True)

if __name__ == "__main__":
    linear_case(98304, 320, 320)
    linear_case(98304, 320, 320, res=True)
    linear_case(98304, 960, 320)
    linear_case(98304, 2560, 320, geglu=True)
    linear_case(98304, 320, 1280, res=True)
    linear_case(24576, 640, 640, res=True)
    linear_case(24576, 5120, 640, geglu=True)
    linear_case(6144, 1280, 1280, res=True)
    linear_case(1848, 640, 768)
    linear_case(1848, 2560, 768)
    linear_case(924, 2560, 768)
    conv_case(24, 64, 64, 320, 320)
    conv_case(24, 32, 32, 640, 640)
    conv_case(24, 16, 16, 1280, 1280)
    conv_case(24, 8, 8, 1280, 1280)
    conv_case(12, 8, 8, 1280, 1280)
