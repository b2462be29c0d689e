"""Which kernels run power-capped?  Each of the step's kernel classes in a ~2.5 s loop with nvidia-smi sampling the SM
clock and board power every 100 ms: the step as a whole settles at ~1.8 GHz under the 1 kW cap (sw_power_cap), and a
kernel's share of the ENERGY, not only of the time, decides how much a speed-up elsewhere is worth."""
import os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from storygen_b200 import ops, _lib
from storygen_b200.ops import BF16, Src, taps3x3

dev = "cuda"


class Smi:
    def __init__(self):
        self.rows = []
        self.p = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits",
                                   "-lms", "100"], stdout=subprocess.PIPE, text=True)
        threading.Thread(target=self._rd, daemon=True).start()

    def _rd(self):
        for l in self.p.stdout:
            try:
                c, w = l.split(",")
                self.rows.append((time.time(), int(c), float(w)))
            except Exception:
                pass

    def window(self, t0, t1):
        r = [x for x in self.rows if t0 + 0.5 <= x[0] <= t1]
        if not r:
            return None, None
        return sorted(x[1] for x in r)[len(r) // 2], sum(x[2] for x in r) / len(r)


def loop(name, fn, flops, smi, secs=2.5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.time()
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < secs:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    ms = e0.elapsed_time(e1) / n
    clk, w = smi.window(t0, t1)
    print(f"{name:44s} {ms * 1e3:9.1f} us  {flops / ms / 1e9:7.0f} TF/s  clk {clk} MHz  {w:.0f} W", flush=True)
    time.sleep(1.0)


def main():
    smi = Smi()
    time.sleep(1.0)
    lanes, c = 24, 320
    q = torch.randn(lanes, 4096, c, device=dev).to(BF16)
    o = torch.empty_like(q)
    loop("attn d=40 24 lanes 4096x4096", lambda: ops.attention(q, q, q, o, lanes, 8, 40, 4096, 4096, (4096 * c, c), (4096 * c, c), (4096 * c, c)),
         4.0 * lanes * 4096 * 4096 * c, smi)
    x = torch.randn(lanes, 64, 64, c, device=dev).to(BF16)
    y = torch.empty_like(x)
    w = (torch.randn(c, 9 * c, device=dev) * (9 * c) ** -0.5).to(BF16)
    loop("conv3x3 320->320 64x64 24 lanes", lambda: ops.gemm([Src.nhwc(x)], taps3x3(0, c), (64, 64, lanes), w, y), 2.0 * lanes * 4096 * c * 9 * c, smi)
    c2 = 1280
    x2 = torch.randn(lanes, 16, 16, c2, device=dev).to(BF16)
    y2 = torch.empty_like(x2)
    w2 = (torch.randn(c2, 9 * c2, device=dev) * (9 * c2) ** -0.5).to(BF16)
    loop("conv3x3 1280->1280 16x16 24 lanes", lambda: ops.gemm([Src.nhwc(x2)], taps3x3(0, c2), (16, 16, lanes), w2, y2), 2.0 * lanes * 256 * c2 * 9 * c2, smi)
    m = lanes * 4096
    xa = torch.randn(m, c, device=dev).to(BF16)
    wf = (torch.randn(8 * c, c, device=dev) * c ** -0.5).to(BF16)
    ff = torch.empty(m, 4 * c, device=dev, dtype=BF16)
    loop("GEGLU FF1 98304x2560x320", lambda: ops.linear(xa, wf, ff, mode=_lib.SG_EPI_GEGLU), 2.0 * m * 8 * c * c, smi)
    wq = (torch.randn(c, c, device=dev) * c ** -0.5).to(BF16)
    yo = torch.empty(m, c, device=dev, dtype=BF16)
    loop("linear 98304x320x320 + residual", lambda: ops.linear(xa, wq, yo, residual=xa), 2.0 * m * c * c, smi)
    st = ops.groupnorm_ws(lanes, 32, dev)
    g, b = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    xo = torch.empty(lanes, 4096, c, device=dev, dtype=BF16)
    loop("groupnorm stats+apply 24x4096x320", lambda: ops.groupnorm(x.view(lanes, 4096, c), None, lanes, 4096, 32, 1e-5, g, b, True, xo, st), 0.0, smi)
    a = torch.empty(8192, 8192, device=dev, dtype=BF16).normal_()
    bm = torch.empty(8192, 8192, device=dev, dtype=BF16).normal_()
    loop("cuBLAS bf16 8192^3 (reference point)", lambda: torch.matmul(a, bm), 2.0 * 8192 ** 3, smi)
    smi.p.terminate()


if __name__ == "__main__":
    main()
