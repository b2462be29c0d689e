"""Torch-tensor front end of the C ABI: every function checks device/dtype/contiguity, builds the
argument struct and enqueues ONE kernel on torch's current CUDA stream.  PyTorch is used only for
device memory and streams; none of these functions has a torch / CPU fallback."""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import SG_EPI_GEGLU, SG_EPI_LINEAR, SG_EPI_NCHW_F32, SgAttnArgs, SgAttnBwdArgs, SgGemmArgs, SgGnParts

BF16 = torch.bfloat16
TRACE = None  # set to a list to record (kernel, shape...) of every launch, in order (profiling aid)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (the StoryGen hot path has no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous tensor")


class Src:
    """One NHWC activation source of the implicit-GEMM A operand: (ptr, dims, strides) in elements."""

    __slots__ = ("t", "ptr", "c", "w", "h", "n", "sw", "sh", "sn")

    def __init__(self, t: torch.Tensor, c: int, w: int, h: int, n: int, sw: int, sh: int, sn: int, offset: int = 0):
        self.t = t  # keeps the storage alive
        self.ptr = t.data_ptr() + 2 * offset
        self.c, self.w, self.h, self.n, self.sw, self.sh, self.sn = c, w, h, n, sw, sh, sn

    @staticmethod
    def rows(t: torch.Tensor) -> "Src":
        """(M, K) row-major matrix as a 1 x 1 x M 'image' with K channels."""
        m, k = t.shape
        return Src(t, k, m, 1, 1, k, m * k, m * k)

    @staticmethod
    def nhwc(t: torch.Tensor) -> "Src":
        n, h, w, c = t.shape
        return Src(t, c, w, h, n, c, w * c, h * w * c)

    @staticmethod
    def nhwc_phase(t: torch.Tensor, py: int, px: int) -> "Src":
        """Every second pixel starting at (py, px): the 4 phases turn a stride-2 conv into stride-1 taps."""
        n, h, w, c = t.shape
        return Src(t, c, w // 2, h // 2, n, 2 * c, 2 * w * c, h * w * c, offset=(py * w + px) * c)


def tile_box(w: int, h: int, n: int) -> Tuple[int, int, int]:
    """128-row tile as a bw x bh x bn box of the output pixel grid.  A plain (M, K) matrix (h == n == 1)
    always uses 128 consecutive rows: TMA clips the ragged last tile, the epilogue masks it."""
    if h == 1 and n == 1:
        return 128, 1, 1
    bw = math.gcd(w, 128)
    bh = math.gcd(h, 128 // bw)
    return bw, bh, 128 // (bw * bh)


_PARTS_CACHE = {}


def row_stat_parts(grid: Tuple[int, int, int], n_cols: int, k_total: int) -> int:
    """Number of row-statistics partials a LINEAR-mode GEMM with this output grid, N and K writes per row
    (sg_gemm_row_stat_parts: 2 x its N tiles, whose width depends on N, K and the tile count)."""
    key = (grid, n_cols, k_total, torch.cuda.current_device())
    v = _PARTS_CACHE.get(key)
    if v is None:
        a = SgGemmArgs()
        a.out_w, a.out_h, a.out_n = grid
        a.bw, a.bh, a.bn = tile_box(*grid)
        a.N, a.Ktot, a.mode = n_cols, k_total, SG_EPI_LINEAR
        v = _lib.load().sg_gemm_row_stat_parts(C.byref(a))
        if v < 0:
            _lib.check(v, "sg_gemm_row_stat_parts")
        _PARTS_CACHE[key] = v
    return v


class ColStats:
    """Per-channel GroupNorm partials of a tensor, as written by the `reps` GEMMs that produced it on the (w, h, n) pixel
    grid (sg_gemm col_stats_out; reps = 4 for the sub-pixel phases of an up-sampling conv).  buf: fp32 (reps, blocks, c, 2)."""
    __slots__ = ("buf", "c", "tiles_xy", "bn", "reps", "m_tiles")

    @staticmethod
    def layout(grid: Tuple[int, int, int]):
        """(m_tiles, tiles_xy, imgs_per_tile) or None when a 32-row quarter would span two images (fewer than 32 pixels
        per image and tile: the consumer then falls back to the separate statistics pass, sg_groupnorm_stats)."""
        w, h, n = grid
        bw, bh, bn = tile_box(w, h, n)
        if (bw * bh) % 32 != 0 or bn > 4:
            return None
        tx, ty, tn = -(-w // bw), -(-h // bh), -(-n // bn)
        return tx * ty * tn, tx * ty, bn

    def __init__(self, grid: Tuple[int, int, int], c: int, device, reps: int = 1):
        self.m_tiles, self.tiles_xy, self.bn = ColStats.layout(grid)
        self.c, self.reps = c, reps
        self.buf = torch.empty(reps, self.m_tiles * 4, c, 2, dtype=torch.float32, device=device)

    def rep(self, i: int) -> torch.Tensor:
        return self.buf[i]

    def desc(self) -> SgGnParts:
        d = SgGnParts()
        d.parts, d.c, d.tiles_xy, d.imgs_per_tile, d.reps, d.m_tiles = self.buf.data_ptr(), self.c, self.tiles_xy, self.bn, self.reps, self.m_tiles
        return d


def gemm(srcs: Sequence[Src], segs: Sequence[Tuple[int, int, int, int, int]], grid: Tuple[int, int, int],
         weight: torch.Tensor, out: torch.Tensor, bias: Optional[torch.Tensor] = None,
         lane_bias: Optional[torch.Tensor] = None, ld_lane_bias: int = 0, residual: Optional[torch.Tensor] = None,
         res_scale: float = 1.0, mode: int = SG_EPI_LINEAR, out_channels: int = 0,
         row_stats_out: Optional[torch.Tensor] = None, ln=None, col_stats_out: Optional[torch.Tensor] = None,
         out_strides: Optional[Tuple[int, int, int]] = None, out_offset: int = 0) -> torch.Tensor:
    """out[m, :] = epilogue(sum_k A[m, k] W[:, k]); see include/storygen_b200.h (sg_gemm).
    segs: (src, dx, dy, c0, nkb); grid: (out_w, out_h, out_n); weight: bf16 [N, Ktot].
    row_stats_out: fp32 (row_stat_parts(grid, N), M, 2) receiving the output's per-row LayerNorm partials;
    ln = (stats (parts, M, 2), csum (N,), eps): apply LayerNorm of the A rows in the epilogue (weight pre-scaled by gamma);
    col_stats_out: fp32 (4 * M tiles, N, 2) receiving the output's per-channel GroupNorm partials (ColStats.rep(i));
    out_strides / out_offset: element strides of (x, y, n) and first element when the output pixels are a strided
    sub-grid of `out` (sub-pixel phase of an up-sampling conv)."""
    a = SgGemmArgs()
    a.num_src = len(srcs)
    for i, s in enumerate(srcs):
        e = a.src[i]
        e.ptr, e.dim_c, e.dim_w, e.dim_h, e.dim_n = s.ptr, s.c, s.w, s.h, s.n
        e.stride_w, e.stride_h, e.stride_n = s.sw, s.sh, s.sn
    a.num_segs = len(segs)
    for i, (src, dx, dy, c0, nkb) in enumerate(segs):
        e = a.segs[i]
        e.src, e.dx, e.dy, e.c0, e.nkb = src, dx, dy, c0, nkb
    a.out_w, a.out_h, a.out_n = grid
    a.bw, a.bh, a.bn = tile_box(*grid)
    _req(weight, BF16, "weight")
    _req(out, torch.float32 if mode == SG_EPI_NCHW_F32 else BF16, "out")
    a.out_channels = out_channels
    a.b = weight.data_ptr()
    a.N, a.Ktot = weight.shape
    a.mode = mode
    a.res_scale = res_scale
    a.bias = _ptr(bias)
    a.lane_bias = _ptr(lane_bias)
    a.ld_lane_bias = ld_lane_bias
    a.residual = _ptr(residual)
    a.ldr = residual.shape[-1] if residual is not None else 0
    a.out = out.data_ptr() + out_offset * out.element_size()
    a.ldo = out.shape[-1]
    if out_strides is not None:
        a.out_stride_w, a.out_stride_h, a.out_stride_n = out_strides
    if row_stats_out is not None:
        _req(row_stats_out, torch.float32, "row_stats_out")
        want = _lib.load().sg_gemm_row_stat_parts(C.byref(a))
        if row_stats_out.shape[0] != want or row_stats_out.shape[1] != grid[0] * grid[1] * grid[2]:
            raise ValueError(f"row_stats_out must be ({want}, M, 2) for this GEMM, got {tuple(row_stats_out.shape)}")
        a.row_stats_out = row_stats_out.data_ptr()
    if col_stats_out is not None:
        _req(col_stats_out, torch.float32, "col_stats_out")
        a.col_stats_out = col_stats_out.data_ptr()
    if ln is not None:
        st, csum, eps = ln
        _req(st, torch.float32, "ln stats")
        _req(csum, torch.float32, "ln csum")
        if st.shape[1] != grid[0] * grid[1] * grid[2] or csum.numel() != weight.shape[0]:
            raise ValueError("ln statistics / csum do not match the GEMM's rows / columns")
        a.ln_stats, a.ln_csum, a.ln_parts, a.ln_eps = st.data_ptr(), csum.data_ptr(), st.shape[0], eps
    if TRACE is not None:
        TRACE.append(("gemm", grid[0] * grid[1] * grid[2], int(a.N), int(a.Ktot), len(segs), mode))
    _lib.check(_lib.load().sg_gemm(C.byref(a), _stream()), "sg_gemm")
    return out


def linear(x: torch.Tensor, weight: torch.Tensor, out: torch.Tensor, bias=None, residual=None, res_scale=1.0,
           mode=SG_EPI_LINEAR, row_stats_out=None, ln=None, col_stats_out=None) -> torch.Tensor:
    """x (M, K) bf16 @ weight (N, K)^T with the fused epilogue."""
    m, k = x.shape
    return gemm([Src.rows(x)], [(0, 0, 0, 0, k // 64)], (m, 1, 1), weight, out, bias=bias, residual=residual,
                res_scale=res_scale, mode=mode, row_stats_out=row_stats_out, ln=ln, col_stats_out=col_stats_out)


def taps3x3(src: int, cin: int, c0: int = 0) -> List[Tuple[int, int, int, int, int]]:
    return [(src, kx - 1, ky - 1, c0, cin // 64) for ky in range(3) for kx in range(3)]


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, n_lanes: int, n_heads: int,
              head_dim: int, nq: int, nk: int, q_strides: Tuple[int, int], k_strides: Tuple[int, int],
              v_strides: Tuple[int, int], n_kv_lanes: int = 0, kv_map: Optional[torch.Tensor] = None,
              lse: Optional[torch.Tensor] = None) -> torch.Tensor:
    """softmax(Q K^T / sqrt(d)) V per (lane, head); q/k/v may be strided views into fused projections.
    *_strides = (lane stride, token stride) in elements.  lse: optional fp32 (lanes, heads, nq) receiving the log2-domain
    log-sum-exp of every score row (saved for attention_bwd)."""
    a = SgAttnArgs()
    a.q, a.k, a.v, a.o = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    a.q_stride_n, a.q_stride_t = q_strides
    a.k_stride_n, a.k_stride_t = k_strides
    a.v_stride_n, a.v_stride_t = v_strides
    c = n_heads * head_dim
    a.o_stride_n, a.o_stride_t = nq * c, c
    a.n_lanes, a.n_heads, a.head_dim, a.nq, a.nk = n_lanes, n_heads, head_dim, nq, nk
    a.n_kv_lanes = n_kv_lanes
    a.kv_map = _ptr(kv_map)
    a.scale = head_dim ** -0.5
    if lse is not None:
        _req(lse, torch.float32, "lse")
        a.lse = lse.data_ptr()
    if TRACE is not None:
        TRACE.append(("attn", n_lanes, n_heads, head_dim, nq, nk))
    _lib.check(_lib.load().sg_attention(C.byref(a), _stream()), "sg_attention")
    return out


ATTN_BWD_HEAD_DIMS = (32, 40, 64, 80, 128)


def attention_bwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, o: torch.Tensor, d_o: torch.Tensor, lse: torch.Tensor,
                  n_heads: int):
    """Gradients of attention(q, k, v) for contiguous bf16 (lanes, tokens, heads * head_dim) tensors, given the forward
    output o, its log-sum-exp rows lse and the output gradient d_o.  Returns (dq, dk, dv) in bf16.
    One sg_attention_bwd call: a row-dot kernel (delta) + the tcgen05 backward kernel."""
    for t, name in ((q, "q"), (k, "k"), (v, "v"), (o, "o"), (d_o, "d_o")):
        _req(t, BF16, name)
    _req(lse, torch.float32, "lse")
    n, nq, c = q.shape
    nk = k.shape[1]
    hd = c // n_heads
    if hd not in ATTN_BWD_HEAD_DIMS:
        raise NotImplementedError(f"attention_bwd: head dim {hd} (supported: {ATTN_BWD_HEAD_DIMS})")
    dq_acc = torch.zeros(n, nq, c, dtype=torch.float32, device=q.device)
    dk, dv = torch.empty_like(k), torch.empty_like(v)
    delta = torch.empty(n, n_heads, nq, dtype=torch.float32, device=q.device)
    a = SgAttnBwdArgs()
    a.q, a.k, a.v, a.o, a.d_o = q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), d_o.data_ptr()
    a.lse, a.delta, a.dq_accum, a.dk, a.dv = lse.data_ptr(), delta.data_ptr(), dq_acc.data_ptr(), dk.data_ptr(), dv.data_ptr()
    a.q_stride_n, a.q_stride_t = nq * c, c
    a.k_stride_n, a.k_stride_t = nk * c, c
    a.v_stride_n, a.v_stride_t = nk * c, c
    a.o_stride_n, a.o_stride_t = nq * c, c
    a.dk_stride_n, a.dk_stride_t = nk * c, c
    a.n_lanes, a.n_heads, a.head_dim, a.nq, a.nk = n, n_heads, hd, nq, nk
    a.scale = hd ** -0.5
    _lib.check(_lib.load().sg_attention_bwd(C.byref(a), _stream()), "sg_attention_bwd")
    return dq_acc.to(BF16), dk, dv


def groupnorm(x0: torch.Tensor, x1: Optional[torch.Tensor], n: int, hw: int, groups: int, eps: float,
              gamma: torch.Tensor, beta: torch.Tensor, silu: bool, out: torch.Tensor, stats: torch.Tensor):
    """y = [SiLU](GroupNorm(concat_c(x0, x1))); x* are (n, hw, c*) bf16.  Two launches (stats, apply)."""
    lib = _lib.load()
    c0 = x0.shape[-1]
    c1 = x1.shape[-1] if x1 is not None else 0
    _lib.check(lib.sg_groupnorm_stats(x0.data_ptr(), c0, _ptr(x1), c1, n, hw, groups, eps, stats.data_ptr(),
                                      _stream()), "sg_groupnorm_stats")
    _lib.check(lib.sg_groupnorm_apply(x0.data_ptr(), c0, _ptr(x1), c1, n, hw, groups, eps, stats.data_ptr(),
                                      gamma.data_ptr(), beta.data_ptr(), int(silu), out.data_ptr(), _stream()),
               "sg_groupnorm_apply")
    return out


def groupnorm_from_parts(x0: torch.Tensor, st0: ColStats, x1: Optional[torch.Tensor], st1: Optional[ColStats], n: int, hw: int,
                         groups: int, eps: float, gamma: torch.Tensor, beta: torch.Tensor, silu: bool, out: torch.Tensor,
                         mean_rstd: torch.Tensor):
    """y = [SiLU](GroupNorm(concat_c(x0, x1))) with the statistics the PRODUCING GEMMs emitted (col_stats_out): one tiny
    reduction kernel (fixed order, fp64) + the apply pass; no pass over the activations for the statistics."""
    lib = _lib.load()
    c0 = x0.shape[-1]
    c1 = x1.shape[-1] if x1 is not None else 0
    d0 = st0.desc()
    d1 = st1.desc() if st1 is not None else None
    _lib.check(lib.sg_groupnorm_finalize(C.byref(d0), C.byref(d1) if d1 is not None else None, n, hw, groups, eps,
                                         mean_rstd.data_ptr(), _stream()), "sg_groupnorm_finalize")
    _lib.check(lib.sg_groupnorm_apply_ms(x0.data_ptr(), c0, _ptr(x1), c1, n, hw, groups, mean_rstd.data_ptr(),
                                         gamma.data_ptr(), beta.data_ptr(), int(silu), out.data_ptr(), _stream()),
               "sg_groupnorm_apply_ms")
    return out


def groupnorm_ws(n: int, groups: int, device) -> torch.Tensor:
    return torch.empty(_lib.load().sg_groupnorm_ws_floats(n, groups), dtype=torch.float32, device=device)


def layernorm(x: torch.Tensor, eps: float, ga, ba, ya, gb=None, bb=None, yb=None):
    rows, c = x.shape
    _lib.check(_lib.load().sg_layernorm(x.data_ptr(), rows, c, eps, ga.data_ptr(), ba.data_ptr(), ya.data_ptr(),
                                        _ptr(gb), _ptr(bb), _ptr(yb), _stream()), "sg_layernorm")


def timestep_features(t: torch.Tensor, dim: int, out: torch.Tensor):
    _req(t, torch.float32, "timesteps")
    _lib.check(_lib.load().sg_timestep_features(t.data_ptr(), t.numel(), dim, out.data_ptr(), _stream()),
               "sg_timestep_features")
    return out


def small_linear(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], out: torch.Tensor, silu_in=False,
                 silu_out=False):
    n, k = x.shape
    _lib.check(_lib.load().sg_small_linear(x.data_ptr(), n, k, w.data_ptr(), _ptr(b), w.shape[0], int(silu_in),
                                           int(silu_out), out.data_ptr(), _stream()), "sg_small_linear")
    return out


def upsample2x(x: torch.Tensor, out: torch.Tensor):
    n, h, w, c = x.shape
    _lib.check(_lib.load().sg_upsample2x(x.data_ptr(), n, h, w, c, out.data_ptr(), _stream()), "sg_upsample2x")
    return out


def im2col3x3_nchw(x_nchw: torch.Tensor, out: torch.Tensor):
    """(n, cin, h, w) fp32 -> (n*h*w, 64) bf16 rows of zero-padded 3x3 patches, column (ky*3+kx)*cin + ci."""
    _req(x_nchw, torch.float32, "sample")
    n, cin, h, ww = x_nchw.shape
    _lib.check(_lib.load().sg_im2col3x3_nchw(x_nchw.data_ptr(), n, cin, h, ww, out.data_ptr(), _stream()),
               "sg_im2col3x3_nchw")
    return out


def add_noise(x: torch.Tensor, eps: torch.Tensor, acp: torch.Tensor, t: torch.Tensor, out: torch.Tensor):
    """out[n] = sqrt(acp[t[n]]) x[n] + sqrt(1 - acp[t[n]]) eps[n]; fp32, t int32 (device)."""
    n = x.shape[0]
    _lib.check(_lib.load().sg_add_noise(x.data_ptr(), eps.data_ptr(), acp.data_ptr(), t.data_ptr(), n,
                                        x.numel() // n, out.data_ptr(), _stream()), "sg_add_noise")
    return out


def cfg_ddim_step(eps3: torch.Tensor, latents: torch.Tensor, s_img: float, s_txt: float, alphas: torch.Tensor):
    """In-place dual-guidance DDIM update of `latents` (b, ...) from eps3 (3b, ...); alphas = device (a_t, a_prev)."""
    b = latents.shape[0]
    _lib.check(_lib.load().sg_cfg_ddim_step(eps3.data_ptr(), latents.data_ptr(), b, latents.numel() // b, s_img,
                                            s_txt, alphas.data_ptr(), _stream()), "sg_cfg_ddim_step")
    return latents
