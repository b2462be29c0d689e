"""A few representative GEMM launches for `ncu --set full` captures.  Usage: python scripts/one_gemm.py <case>
cases: qkv (98304 x 960 x 320), geglu (98304 x 2560 x 320, GEGLU), res (98304 x 320 x 320 + residual), conv (3x3 320->320, 24 lanes)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from storygen_b200 import ops, _lib
from storygen_b200.ops import BF16, Src, taps3x3

case = sys.argv[1] if len(sys.argv) > 1 else "qkv"
dev = "cuda"
if case == "conv":
    x = torch.randn(24, 64, 64, 320, device=dev).to(BF16)
    w = (torch.randn(320, 2880, device=dev) * 2880 ** -0.5).to(BF16)
    y = torch.empty(24, 64, 64, 320, device=dev, dtype=BF16)
    fn = lambda: ops.gemm([Src.nhwc(x)], taps3x3(0, 320), (64, 64, 24), w, y, bias=torch.zeros(320, device=dev))
else:
    m, k = 98304, 320
    n = {"qkv": 960, "geglu": 2560, "res": 320}[case]
    x = torch.randn(m, k, device=dev).to(BF16)
    w = (torch.randn(n, k, device=dev) * k ** -0.5).to(BF16)
    b = torch.zeros(n, device=dev)
    nout = n // 2 if case == "geglu" else n
    y = torch.empty(m, nout, device=dev, dtype=BF16)
    r = torch.randn(m, nout, device=dev).to(BF16) if case == "res" else None
    mode = _lib.SG_EPI_GEGLU if case == "geglu" else _lib.SG_EPI_LINEAR
    fn = lambda: ops.linear(x, w, y, bias=b, residual=r, mode=mode)
for _ in range(4):
    fn()
torch.cuda.synchronize()
