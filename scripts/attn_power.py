"""Energy per launch of the d=40 attention core (12 lanes x 8 heads x 4096 x 12288): duration x board power in a 3 s loop.
The step runs power-capped (sw_power_cap), so the joules a kernel needs, not only its isolated duration, decide the step time.
Usage: [SG_ATTN_MC=0] [STORYGEN_B200_LIB=<alt build>] python scripts/attn_power.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time
import torch
from storygen_b200 import ops
from storygen_b200.ops import BF16
from scripts.power_probe import Smi

dev = "cuda"


def main():
    smi = Smi()
    time.sleep(1.0)
    lanes, heads, d, nq, nk = 12, 8, 40, 4096, 12288
    c = heads * d
    q = torch.randn(lanes, nq, c, device=dev).to(BF16)
    k = torch.randn(lanes, nk, c, device=dev).to(BF16)
    v = torch.randn(lanes, nk, c, device=dev).to(BF16)
    o = torch.empty_like(q)
    fn = lambda: ops.attention(q, k, v, o, lanes, heads, d, nq, nk, (nq * c, c), (nk * c, c), (nk * c, c))
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.time()
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < 3.0:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    us = e0.elapsed_time(e1) / n * 1e3
    clk, w = smi.window(t0, t1)
    print(f"{os.environ.get('VARIANT', '?'):12s} {us:9.1f} us  {4.0 * lanes * nq * nk * c / us / 1e6:6.0f} TF/s  clk {clk} MHz  {w:.0f} W  "
          f"{us * w * 1e-3:7.1f} mJ per launch", flush=True)


if __name__ == "__main__":
    main()
