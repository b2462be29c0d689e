"""CUDA-event timing of the attention shapes of one denoising step (B=4, R=3).  Usage: python scripts/attn_bench.py
(STORYGEN_B200_LIB=<path> selects an alternative build of the library for A/B runs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from storygen_b200 import ops
from storygen_b200.ops import BF16

dev = "cuda"


def case(lanes, heads, d, nq, nk, iters=10):
    c = heads * d
    q = torch.randn(lanes, nq, c, device=dev).to(BF16)
    k = torch.randn(lanes, nk, c, device=dev).to(BF16)
    v = torch.randn(lanes, nk, c, device=dev).to(BF16)
    o = torch.empty(lanes, nq, c, device=dev, dtype=BF16)
    fn = lambda: ops.attention(q, k, v, o, lanes, heads, d, nq, nk, (nq * c, c), (nk * c, c), (nk * c, c))
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    fl = 4.0 * lanes * nq * nk * c
    tiles = lanes * heads * ((nq + 127) // 128) * ((nk + 127) // 128)
    print(f"attn lanes={lanes:3d} heads={heads} d={d:3d} nq={nq:5d} nk={nk:6d}: {us:9.1f} us  {fl/us/1e6:6.0f} TF/s  "
          f"{us * 1e-6 * 1.96e9 * 148 / tiles:7.0f} SM-clk per 128x128 tile", flush=True)


if __name__ == "__main__":
    case(24, 8, 40, 4096, 4096)
    case(12, 8, 40, 4096, 4096)
    case(12, 8, 40, 4096, 12288)
    case(24, 8, 80, 1024, 1024)
    case(12, 8, 80, 1024, 3072)
    case(24, 8, 160, 256, 256)
    case(12, 8, 160, 256, 768)
    case(24, 8, 40, 4096, 77)
    case(24, 8, 80, 1024, 77)
