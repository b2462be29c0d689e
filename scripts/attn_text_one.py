"""One 77-token cross-attention launch (24 lanes, d=40) for `ncu --set full` captures.  Usage: python scripts/attn_text_one.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from storygen_b200 import ops
from storygen_b200.ops import BF16

lanes, heads, d, nq, nk = 24, 8, 40, 4096, 77
c = heads * d
q = torch.randn(lanes, nq, c, device="cuda").to(BF16)
k = torch.randn(lanes, nk, c, device="cuda").to(BF16)
v = torch.randn(lanes, nk, c, device="cuda").to(BF16)
o = torch.empty(lanes, nq, c, device="cuda", dtype=BF16)
for _ in range(4):
    ops.attention(q, k, v, o, lanes, heads, d, nq, nk, (nq * c, c), (nk * c, c), (nk * c, c))
torch.cuda.synchronize()
