"""One GroupNorm(+SiLU) of a 24 x 64 x 64 x 320 activation with statistics from the producing GEMM's epilogue, for
`ncu --set full -k regex:gn_apply_ms` / `-k regex:gn_finalize` captures.  Usage: python scripts/gn_one.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from storygen_b200 import ops
from storygen_b200.ops import BF16, Src

n, h, w, c = 24, 64, 64, 320
x = torch.randn(n, h, w, c, device="cuda").to(BF16)
wt = (torch.randn(c, c, device="cuda") * c ** -0.5).to(BF16)
y = torch.empty_like(x)
st = ops.ColStats((w, h, n), c, x.device)
out = torch.empty_like(x)
gamma, beta = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
mr = ops.groupnorm_ws(n, 32, x.device)
for _ in range(4):
    ops.gemm([Src.nhwc(x)], [(0, 0, 0, 0, c // 64)], (w, h, n), wt, y, col_stats_out=st.rep(0))
    ops.groupnorm_from_parts(y, st, None, None, n, h * w, 32, 1e-5, gamma, beta, True, out, mr)
torch.cuda.synchronize()
ref = torch.nn.functional.silu(torch.nn.functional.group_norm(y.float().permute(0, 3, 1, 2), 32, eps=1e-5)).permute(0, 2, 3, 1)
print("max err", (out.float() - ref).abs().max().item())
