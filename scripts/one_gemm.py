import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from storygen_b200 import ops
from storygen_b200.ops import BF16
m, n, k = 98304, 2560, 320
x = torch.randn(m, k, device="cuda").to(BF16); w = (torch.randn(n, k, device="cuda") * k ** -0.5).to(BF16)
b = torch.zeros(n, device="cuda"); y = torch.empty(m, n, device="cuda", dtype=BF16)
for _ in range(3): ops.linear(x, w, y, bias=b)
torch.cuda.synchronize()
