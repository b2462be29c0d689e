"""Executes UNet2DConditionModel.forward with the sm_100a kernels.

Data layout: every activation is channel-last bf16 — images (n, y, x, c), token streams (n, token, c);
the two are the same memory, so the reference's NCHW<->token permutes around each transformer
(model/attention.py:103,117) do not exist here.  Weights are re-packed once per model (bf16, K-major,
conv taps / shortcut / fused projections concatenated along K or N as the kernels consume them).

Normalisation is fused into the GEMMs either side of it:
  * LayerNorm never runs as a kernel: LN(x) W^T + b = rstd (x W'^T - mean csum) + b' with W' = W diag(gamma),
    csum = row sums of W', b' = b + W beta (packed once); the GEMM that PRODUCES x emits per-row (sum, sumsq)
    partials from its epilogue, the consuming GEMM multiplies the raw x and applies mean / rstd per row in its epilogue.
  * GroupNorm statistics come from the producing GEMM's epilogue as per-channel partial sums (32-row blocks), reduced
    by a tiny fixed-order kernel; only the apply(+SiLU, + channel concat) pass touches the activations.

Launch sequence per ResnetBlock2D (6 kernels):
    gn_finalize, gn_apply+SiLU(+channel concat) -> conv3x3 GEMM (+bias +time embedding, emits GN partials)
    gn_finalize, gn_apply+SiLU                  -> conv3x3 GEMM (+1x1 shortcut as extra K segments | +x, emits GN partials)
per Transformer2DModel (feature pass 12, denoise pass 15 kernels; + one text K/V GEMM per forward or per prompt):
    GN -> proj_in GEMM -> QKV GEMM (LN1 folded) -> attention -> out-proj GEMM(+h) [= the tapped feature]
    Q GEMM (LN2 folded), attention over the pre-projected text K/V [-> Q3 GEMM (LN4 folded), image KV GEMM, attention]
    out-proj GEMM over [a_text | a_image] (+2h) -> FF1 GEMM (LN3 folded, GEGLU) -> FF2 GEMM(+h) -> proj_out GEMM(+x)
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from ._lib import SG_EPI_GEGLU, SG_EPI_NCHW_F32
from .ops import BF16, Src, taps3x3

F32 = torch.float32
# SG_FUSE_Q23=0: separate to_q GEMMs for attn2 and attn3 in the denoise pass (A/B timing)
_FUSE_Q23 = os.environ.get("SG_FUSE_Q23", "1") != "0"


class Act:
    """An NHWC activation (n, h, w, c) plus, when its producer(s) emitted them, its per-channel GroupNorm partials
    (ops.ColStats)."""
    __slots__ = ("t", "cs")

    def __init__(self, t, cs=None):
        self.t, self.cs = t, cs

    def stats(self, i=0):
        return None if self.cs is None else self.cs.rep(i)


class UNetEngine:
    def __init__(self, model):
        self.model = model
        p = next(model.parameters())
        if p.device.type != "cuda":
            raise RuntimeError("storygen_b200.UNet2DConditionModel runs only on a CUDA (sm_100a) device: move the "
                               "model with .to('cuda') first (there is no CPU fallback)")
        self.dev = p.device
        cfg = model.config
        self.groups = cfg.norm_num_groups
        self.eps = cfg.norm_eps
        self.heads = cfg.attention_head_dim
        self.boc = tuple(cfg.block_out_channels)
        self.temb_dim = self.boc[0] * 4
        for c in self.boc:
            if c % 64 or (c // self.heads) not in (32, 40, 64, 80, 128, 160):
                raise NotImplementedError(f"block_out_channels {self.boc} / heads {self.heads}: channel counts must be "
                                          "multiples of 64 with head dim in {32,40,64,80,128,160}")
        if cfg.cross_attention_dim % 64:
            raise NotImplementedError("cross_attention_dim must be a multiple of 64")
        self._pack()
        self.fuse_gn_stats = True  # GroupNorm statistics from the producing GEMM's epilogue (False: separate pass)
        self._stats = None      # workspace of the forward in flight
        self._own_stats = None  # engine-owned workspace for direct unet(...) calls (never handed to a captured graph)

    # ------------------------------------------------------------------------------------------
    # weight packing
    # ------------------------------------------------------------------------------------------
    def _bf(self, t):
        return t.detach().to(self.dev, BF16).contiguous()

    def _f32(self, t):
        return t.detach().to(self.dev, F32).contiguous()

    def _conv3(self, w):  # [co, ci, 3, 3] -> [co, (ky, kx, ci)]
        return w.detach().permute(0, 2, 3, 1).reshape(w.shape[0], -1)

    def _pack_resnet(self, r, temb_w: List, temb_b: List):
        d = dict(cin=r.in_channels, cout=r.out_channels)
        d["gn1"] = (self._f32(r.norm1.weight), self._f32(r.norm1.bias))
        d["gn2"] = (self._f32(r.norm2.weight), self._f32(r.norm2.bias))
        d["w1"] = self._bf(self._conv3(r.conv1.weight))
        d["temb_off"] = sum(w.shape[0] for w in temb_w)
        temb_w.append(r.time_emb_proj.weight.detach())
        temb_b.append(r.time_emb_proj.bias.detach() + r.conv1.bias.detach())  # conv1 bias folded in
        w2 = self._conv3(r.conv2.weight)
        b2 = r.conv2.bias.detach()
        if r.conv_shortcut is not None:
            w2 = torch.cat([w2, r.conv_shortcut.weight.detach().reshape(r.out_channels, r.in_channels)], dim=1)
            b2 = b2 + r.conv_shortcut.bias.detach()
        d["w2"], d["b2"], d["shortcut"] = self._bf(w2), self._f32(b2), r.conv_shortcut is not None
        return d

    def _pack_upsample(self, conv):
        """Weights of the 4 sub-pixel phases of (nearest 2x up-sample -> conv3x3): along each axis phase 0 sees the input
        pixels (i-1, i) with taps (k0, k1 + k2), phase 1 the pixels (i, i+1) with taps (k0 + k1, k2).  Summed in fp32,
        rounded to bf16 once; packed [co, (ty, tx, ci)]."""
        w = conv.weight.detach().to(self.dev, F32)  # [co, ci, ky, kx]
        comb = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}
        ws = []
        for py in range(2):
            for px in range(2):
                taps = []
                for ty in range(2):
                    for tx in range(2):
                        acc = 0
                        for ky in comb[py][ty]:
                            for kx in comb[px][tx]:
                                acc = acc + w[:, :, ky, kx]
                        taps.append(acc)  # [co, ci]
                ws.append(torch.stack(taps, 1).reshape(w.shape[0], -1).to(BF16).contiguous())
        return dict(w=ws, b=self._f32(conv.bias))

    def _fold_ln(self, w, b, norm, perm=None):
        """LayerNorm folded into the Linear that follows it (model/attention.py:250,268,283,298):
        LN(x) W^T + b = rstd (x W'^T - mean csum) + b';  returns (W' bf16, csum fp32, b' fp32)."""
        w = w.detach().to(self.dev, F32)
        g, beta = norm.weight.detach().to(self.dev, F32), norm.bias.detach().to(self.dev, F32)
        wp = (w * g[None, :]).to(BF16)
        csum = wp.to(F32).sum(1)  # of the ROUNDED weights the tensor core multiplies
        bp = w @ beta
        if b is not None:
            bp = bp + b.detach().to(self.dev, F32)
        if perm is not None:
            wp, csum, bp = perm(wp), perm(csum), perm(bp)
        return wp.contiguous(), csum.contiguous(), bp.contiguous()

    def _pack_transformer(self, t):
        blk = t.transformer_blocks[0]
        c = t.proj_in.weight.shape[0]
        d = dict(c=c, heads=t.num_attention_heads, hd=t.attention_head_dim)
        d["gn"] = (self._f32(t.norm.weight), self._f32(t.norm.bias))
        d["w_in"], d["b_in"] = self._bf(t.proj_in.weight.reshape(c, -1)), self._f32(t.proj_in.bias)
        d["w_out"], d["b_out"] = self._bf(t.proj_out.weight.reshape(t.proj_out.weight.shape[0], -1)), self._f32(t.proj_out.bias)
        a1, a2, a3 = blk.attn1, blk.attn2, blk.attn3
        d["qkv1"] = self._fold_ln(torch.cat([a1.to_q.weight, a1.to_k.weight, a1.to_v.weight], 0), None, blk.norm1)
        d["w_o1"], d["b_o1"] = self._bf(a1.to_out[0].weight), self._f32(a1.to_out[0].bias)
        d["q2"] = self._fold_ln(a2.to_q.weight, None, blk.norm2)
        d["w_kv2"] = torch.cat([a2.to_k.weight.detach(), a2.to_v.weight.detach()], 0)  # fused across layers in _pack
        d["w_o2"], d["b_o2"] = self._bf(a2.to_out[0].weight), self._f32(a2.to_out[0].bias)
        d["q3"] = self._fold_ln(a3.to_q.weight, None, blk.norm4)
        # denoise pass: to_q of attn2 and attn3 read the same rows (LN2 / LN4 of the same residual stream share their
        # statistics): ONE GEMM over N = [q2 | q3]
        d["q23"] = tuple(torch.cat([d["q2"][i], d["q3"][i]], 0).contiguous() for i in range(3))
        d["w_kv3"] = self._bf(torch.cat([a3.to_k.weight, a3.to_v.weight], 0))
        d["w_o23"] = self._bf(torch.cat([a2.to_out[0].weight, a3.to_out[0].weight], 1))
        d["b_o23"] = self._f32(a2.to_out[0].bias.detach() + a3.to_out[0].bias.detach())
        # GEGLU: interleave [16 value rows | 16 gate rows] so one 32-column epilogue chunk holds both halves
        w, b = blk.ff.net[0].proj.weight.detach(), blk.ff.net[0].proj.bias.detach()
        half = w.shape[0] // 2

        def interleave(v):  # rows [value | gate] -> blocks of [16 value rows | 16 gate rows]
            tail = v.shape[1:]
            return torch.stack([v[:half].reshape(-1, 16, *tail), v[half:].reshape(-1, 16, *tail)], 1).reshape(2 * half, *tail)

        d["ff1"] = self._fold_ln(w, b, blk.norm3, perm=interleave)
        d["w_ff2"], d["b_ff2"] = self._bf(blk.ff.net[2].weight), self._f32(blk.ff.net[2].bias)
        return d

    def _pack(self):
        m = self.model
        tw, tb = [], []
        ci = m.conv_in.weight.shape[1]
        if 9 * ci > 64:
            raise NotImplementedError("in_channels > 7")
        w_in = torch.zeros(m.conv_in.weight.shape[0], 64, dtype=F32, device=self.dev)
        w_in[:, :9 * ci] = self._conv3(m.conv_in.weight).to(self.dev, F32)  # column (ky*3+kx)*cin + ci
        self.conv_in = (self._bf(w_in), self._f32(m.conv_in.bias))
        te = m.time_embedding
        self.time = (self._bf(te.linear_1.weight), self._f32(te.linear_1.bias), self._bf(te.linear_2.weight),
                     self._f32(te.linear_2.bias))
        self.down = []
        for blk in m.down_blocks:
            d = dict(res=[self._pack_resnet(r, tw, tb) for r in blk.resnets])
            d["attn"] = [self._pack_transformer(t) for t in blk.attentions] if blk.has_cross_attention else None
            if blk.downsamplers is not None:
                cv = blk.downsamplers[0].conv
                d["down"] = (self._bf(self._conv3(cv.weight)), self._f32(cv.bias))
            else:
                d["down"] = None
            self.down.append(d)
        mb = m.mid_block
        self.mid = dict(res=[self._pack_resnet(r, tw, tb) for r in mb.resnets],
                        attn=[self._pack_transformer(t) for t in mb.attentions])
        self.up = []
        for blk in m.up_blocks:
            d = dict(res=[self._pack_resnet(r, tw, tb) for r in blk.resnets])
            d["attn"] = [self._pack_transformer(t) for t in blk.attentions] if blk.has_cross_attention else None
            if blk.upsamplers is not None:
                d["up"] = self._pack_upsample(blk.upsamplers[0].conv)
            else:
                d["up"] = None
            self.up.append(d)
        # attn2 to_k / to_v of all transformers as ONE weight [sum 2c, cross_attention_dim]: the text tokens do not
        # change across denoising steps, so the pipeline projects them once per prompt (project_text)
        order = [t for b in self.down if b["attn"] for t in b["attn"]] + self.mid["attn"] + \
                [t for b in self.up if b["attn"] for t in b["attn"]]
        # key of the last tapped feature in execution order (forward(taps_only=True) stops there)
        self.last_tap_key = None
        for i, b in enumerate(self.up):
            if b["attn"]:
                self.last_tap_key = f"up_{i}_{len(b['attn'])}"
        off = 0
        for t in order:
            t["kv2_off"] = off
            off += t["w_kv2"].shape[0]
        self.w_kv2_all = self._bf(torch.cat([t.pop("w_kv2") for t in order], 0))
        self.n_kv2 = off
        self.temb_w = self._bf(torch.cat(tw, 0))
        self.temb_b = self._f32(torch.cat(tb, 0))
        self.norm_out = (self._f32(m.conv_norm_out.weight), self._f32(m.conv_norm_out.bias))
        co = m.conv_out.weight.shape[0]
        if co > 32:
            raise NotImplementedError("out_channels > 32")
        w_out = torch.zeros(32, 9 * m.conv_out.weight.shape[1], dtype=F32, device=self.dev)
        w_out[:co] = self._conv3(m.conv_out.weight).to(self.dev, F32)
        b_out = torch.zeros(32, dtype=F32, device=self.dev)
        b_out[:co] = m.conv_out.bias.detach().to(self.dev, F32)
        self.conv_out = (self._bf(w_out), b_out, co)

    # ------------------------------------------------------------------------------------------
    # building blocks
    # ------------------------------------------------------------------------------------------
    def _new(self, *shape, dtype=BF16):
        return torch.empty(shape, dtype=dtype, device=self.dev)

    def _new_act(self, n, h, w, c) -> Act:
        """Output buffer of a GEMM on the (w, h, n) pixel grid + the buffer for its GroupNorm partials."""
        ok = self.fuse_gn_stats and ops.ColStats.layout((w, h, n)) is not None
        return Act(self._new(n, h, w, c), ops.ColStats((w, h, n), c, self.dev) if ok else None)

    def _gn(self, a0: Act, a1, gamma_beta, eps, silu):
        """[SiLU](GroupNorm(concat_c(a0, a1))) -> (n, hw, c) tensor."""
        n, h, w, c0 = a0.t.shape
        hw = h * w
        c = c0 + (a1.t.shape[-1] if a1 is not None else 0)
        out = self._new(n, hw, c)
        x0 = a0.t.view(n, hw, -1)
        x1 = None if a1 is None else a1.t.view(n, hw, -1)
        if a0.cs is not None and (a1 is None or a1.cs is not None):
            ops.groupnorm_from_parts(x0, a0.cs, x1, None if a1 is None else a1.cs, n, hw, self.groups, eps,
                                     gamma_beta[0], gamma_beta[1], silu, out, self._new(n, self.groups, 2, dtype=F32))
        else:  # a producer without statistics (fewer than 32 pixels per image and tile): separate statistics pass
            ops.groupnorm(x0, x1, n, hw, self.groups, eps, gamma_beta[0], gamma_beta[1], silu, out, self._stats)
        return out

    def _resnet(self, p, a0: Act, a1, lane_bias) -> Act:
        """x = concat_c(a0, a1).  Returns Act (n, h, w, cout)."""
        n, h, w, _ = a0.t.shape
        cout = p["cout"]
        xn = self._gn(a0, a1, p["gn1"], self.eps, True)
        h1 = self._new_act(n, h, w, cout)
        ops.gemm([Src.nhwc(xn.view(n, h, w, -1))], taps3x3(0, p["cin"]), (w, h, n), p["w1"], h1.t,
                 lane_bias=lane_bias[:, p["temb_off"]:], ld_lane_bias=lane_bias.shape[1], col_stats_out=h1.stats())
        hn = self._gn(h1, None, p["gn2"], self.eps, True)
        out = self._new_act(n, h, w, cout)
        srcs = [Src.nhwc(hn.view(n, h, w, cout))]
        segs = taps3x3(0, cout)
        residual = None
        if p["shortcut"]:
            srcs.append(Src.nhwc(a0.t))
            segs.append((1, 0, 0, 0, a0.t.shape[-1] // 64))
            if a1 is not None:
                srcs.append(Src.nhwc(a1.t))
                segs.append((2, 0, 0, 0, a1.t.shape[-1] // 64))
        else:
            residual = a0.t
        ops.gemm(srcs, segs, (w, h, n), p["w2"], out.t, bias=p["b2"], residual=residual, col_stats_out=out.stats())
        return out

    def project_text(self, encoder_hidden_states, out=None):
        """(n, L, D) text tokens -> (n * L, n_kv2) bf16: attn2 K | V of every transformer, layer l at columns
        [kv2_off, kv2_off + 2c)  (model/attention.py:262 `to_k`/`to_v` of attn2, hoisted out of the step loop)."""
        text = encoder_hidden_states.to(self.dev, BF16).contiguous()
        n, L, d = text.shape
        if out is None:
            out = self._new(n * L, self.n_kv2)
        with torch.cuda.device(self.dev):
            return ops.linear(text.view(n * L, d), self.w_kv2_all, out)

    def _transformer(self, p, xa: Act, text_kv, n_text_tokens, image, kv_map, n_kv_lanes, tap_only=False):
        """xa (n, h, w, c); text_kv (n * L, n_kv2) bf16 from project_text; image None or (n_kv, R*hw, c) bf16.
        Returns (out Act (n, h, w, c), tap (n, hw, c)); tap_only: stop once the tapped feature exists -> (None, tap)."""
        n, h, w, _ = xa.t.shape
        hw, c, heads, hd = h * w, p["c"], p["heads"], p["hd"]
        m = n * hw
        x = xa.t
        xn = self._gn(xa, None, p["gn"], 1e-6, False)

        def row_stats(ncols, ktot):
            return torch.empty(ops.row_stat_parts((m, 1, 1), ncols, ktot), m, 2, dtype=F32, device=self.dev)

        h0, rs0 = self._new(m, c), row_stats(c, c)
        ops.linear(xn.view(m, c), p["w_in"], h0, bias=p["b_in"], row_stats_out=rs0)
        # --- self attention (LN1 folded into the QKV projection)
        qkv = self._new(m, 3 * c)
        ops.linear(h0, p["qkv1"][0], qkv, bias=p["qkv1"][2], ln=(rs0, p["qkv1"][1], 1e-5))
        a1 = self._new(m, c)
        st = (hw * 3 * c, 3 * c)
        ops.attention(qkv, qkv[:, c:], qkv[:, 2 * c:], a1, n, heads, hd, hw, hw, st, st, st)
        h1, rs1 = self._new(m, c), row_stats(c, c)  # residual stream after self-attention == the tapped diffusion feature
        ops.linear(a1, p["w_o1"], h1, bias=p["b_o1"], residual=h0, row_stats_out=rs1)
        if tap_only:
            return None, h1.view(n, hw, c)
        # --- text (+ image) cross attention (LN2 / LN4 folded into the Q projections; one GEMM for both when attn3 runs)
        if image is not None and _FUSE_Q23:
            q23 = self._new(m, 2 * c)
            ops.linear(h1, p["q23"][0], q23, bias=p["q23"][2], ln=(rs1, p["q23"][1], 1e-5))
            q2, q3, ldq = q23, q23[:, c:], 2 * c
        else:
            q2, ldq = self._new(m, c), c
            ops.linear(h1, p["q2"][0], q2, bias=p["q2"][2], ln=(rs1, p["q2"][1], 1e-5))
            if image is not None:
                q3 = self._new(m, c)
                ops.linear(h1, p["q3"][0], q3, bias=p["q3"][2], ln=(rs1, p["q3"][1], 1e-5))
        L = n_text_tokens
        ld = text_kv.shape[1]
        kv2 = text_kv[:, p["kv2_off"]:]
        a2 = self._new(m, c)
        ops.attention(q2, kv2, kv2[:, c:], a2, n, heads, hd, hw, L, (hw * ldq, ldq), (L * ld, ld), (L * ld, ld))
        h2, rs2 = self._new(m, c), row_stats(c, 2 * c if image is not None else c)
        if image is not None:
            nkv, nk = image.shape[0], image.shape[1]
            kv3 = self._new(nkv * nk, 2 * c)
            ops.linear(image.view(nkv * nk, c), p["w_kv3"], kv3)
            a3 = self._new(m, c)
            ops.attention(q3, kv3, kv3[:, c:], a3, n, heads, hd, hw, nk, (hw * ldq, ldq), (nk * 2 * c, 2 * c),
                          (nk * 2 * c, 2 * c), n_kv_lanes=nkv, kv_map=kv_map)
            # h_t + h_i = 2h + a_t + a_i  (model/attention.py:277,291,293): one GEMM over K = [a2 | a3]
            ops.gemm([Src.rows(a2), Src.rows(a3)], [(0, 0, 0, 0, c // 64), (1, 0, 0, 0, c // 64)], (m, 1, 1),
                     p["w_o23"], h2, bias=p["b_o23"], residual=h1, res_scale=2.0, row_stats_out=rs2)
        else:
            ops.linear(a2, p["w_o2"], h2, bias=p["b_o2"], residual=h1, row_stats_out=rs2)
        # --- feed forward (LN3 folded into the GEGLU projection)
        ff = self._new(m, 4 * c)
        ops.linear(h2, p["ff1"][0], ff, bias=p["ff1"][2], mode=SG_EPI_GEGLU, ln=(rs2, p["ff1"][1], 1e-5))
        h3 = self._new(m, c)
        ops.linear(ff, p["w_ff2"], h3, bias=p["b_ff2"], residual=h2)
        # proj_out is a (m, 1, 1)-grid GEMM whose rows ARE the pixels of the (w, h, n) grid; its GroupNorm partials must
        # follow that grid's tiling, so it runs on the pixel grid (1x1 conv form)
        out = self._new_act(n, h, w, x.shape[-1])
        ops.gemm([Src.nhwc(h3.view(n, h, w, c))], [(0, 0, 0, 0, c // 64)], (w, h, n), p["w_out"], out.t, bias=p["b_out"],
                 residual=x.view(m, -1), col_stats_out=out.stats())
        return out, h1.view(n, hw, c)

    def _downsample(self, wb, xa: Act) -> Act:
        x = xa.t
        n, h, w, c = x.shape
        out = self._new_act(n, h // 2, w // 2, c)
        srcs = [Src.nhwc_phase(x, py, px) for py in range(2) for px in range(2)]
        # input row 2*oy + ky - 1: ky=0 -> odd phase shifted by -1, ky=1 -> even phase, ky=2 -> odd phase
        ph, sh = (1, 0, 1), (-1, 0, 0)
        segs = [(ph[ky] * 2 + ph[kx], sh[kx], sh[ky], 0, c // 64) for ky in range(3) for kx in range(3)]
        ops.gemm(srcs, segs, (w // 2, h // 2, n), wb[0], out.t, bias=wb[1], col_stats_out=out.stats())
        return out

    def _upsample(self, wp, xa: Act) -> Act:
        """Upsample2D (model/unet_2d_blocks.py:581-584: nearest 2x, then conv3x3) as four 2x2-tap convolutions of the
        LOW-resolution tensor, one per output sub-pixel phase (py, px), with the 3x3 taps that fall on the same input
        pixel pre-summed (_pack_upsample): 4/9 of the FLOPs, the up-sampled tensor is never materialised, and every
        phase writes every second pixel of the output through the GEMM's strided store."""
        x = xa.t
        n, h, w, c = x.shape
        H, W = 2 * h, 2 * w
        ok = self.fuse_gn_stats and ops.ColStats.layout((w, h, n)) is not None
        out = Act(self._new(n, H, W, c), ops.ColStats((w, h, n), c, self.dev, reps=4) if ok else None)
        for py in range(2):
            for px in range(2):
                # phase 0 reads input rows (y-1, y), phase 1 rows (y, y+1); same along x
                segs = [(0, px - 1 + tx, py - 1 + ty, 0, c // 64) for ty in range(2) for tx in range(2)]
                ops.gemm([Src.nhwc(x)], segs, (w, h, n), wp["w"][py * 2 + px], out.t, bias=wp["b"],
                         col_stats_out=out.stats(py * 2 + px), out_strides=(2 * c, 2 * W * c, H * W * c),
                         out_offset=(py * W + px) * c)
        return out

    # ------------------------------------------------------------------------------------------
    # forward
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, sample, timestep, encoder_hidden_states, image_hidden_states=None, kv_map=None, text_kv=None,
                stats=None, taps_only=False):
        """text_kv: optional result of project_text(encoder_hidden_states) (cached across steps by the pipeline);
        stats: optional caller-owned GroupNorm workspace (ops.groupnorm_ws) — a captured CUDA graph must own its own;
        taps_only: the caller only consumes the tapped features (the pipeline's feature pass, model/pipeline.py:431-436
        keeps `[1]` of the UNet output and drops the sample): the forward stops after the last tap — the rest of the last
        transformer block, conv_norm_out and conv_out of that pass are dead code — and returns (None, taps)."""
        with torch.cuda.device(self.dev):  # every launcher enqueues on the current device's current stream
            return self._forward(sample, timestep, encoder_hidden_states, image_hidden_states, kv_map, text_kv, stats,
                                 taps_only)

    def _forward(self, sample, timestep, encoder_hidden_states, image_hidden_states, kv_map, text_kv, stats,
                 taps_only=False):
        dev = self.dev
        n, _, h, w = sample.shape
        if (h % (1 << self.model.num_upsamplers)) or (w % (1 << self.model.num_upsamplers)):
            raise NotImplementedError("latent height/width must be divisible by 2**num_upsamplers")
        sample = sample.to(dev, F32).contiguous()
        if torch.is_tensor(timestep):
            t = timestep.to(dev, F32).reshape(-1)
            t = t.expand(n).contiguous() if t.numel() == 1 else t.contiguous()
        else:
            t = torch.full((n,), float(timestep), device=dev, dtype=F32)
        if t.numel() != n:
            raise ValueError(f"timestep has {t.numel()} entries for a batch of {n}")
        text3 = encoder_hidden_states.to(dev, BF16).contiguous()
        if text3.shape[0] != n or text3.shape[2] != self.model.config.cross_attention_dim:
            raise ValueError(f"encoder_hidden_states {tuple(text3.shape)} does not match batch {n} / "
                             f"cross_attention_dim {self.model.config.cross_attention_dim}")
        L = text3.shape[1]
        if text_kv is None:
            text_kv = self.project_text(text3)
        elif tuple(text_kv.shape) != (n * L, self.n_kv2):
            raise ValueError(f"text_kv has shape {tuple(text_kv.shape)}, expected {(n * L, self.n_kv2)}")
        img = None
        if image_hidden_states is not None:
            img = {k: v.to(dev, BF16).contiguous() for k, v in image_hidden_states.items()}
        if kv_map is not None:
            kv_map = kv_map.to(dev, torch.int32).contiguous()
        need = ops._lib.load().sg_groupnorm_ws_floats(n, self.groups)
        if stats is not None:
            if stats.numel() < need or stats.device != dev or stats.dtype != F32:
                raise ValueError(f"stats workspace needs {need} fp32 elements on {dev}")
            self._stats = stats
        else:
            if self._own_stats is None or self._own_stats.numel() < need:
                self._own_stats = torch.empty(need, dtype=F32, device=dev)
            self._stats = self._own_stats

        # time embedding: features -> linear_1 -> SiLU -> linear_2, then all 22 time_emb_proj at once
        boc0 = self.boc[0]
        tf = ops.timestep_features(t, boc0, self._new(n, boc0, dtype=F32))
        e1 = ops.small_linear(tf, self.time[0], self.time[1], self._new(n, self.temb_dim, dtype=F32), silu_out=True)
        emb = ops.small_linear(e1, self.time[2], self.time[3], self._new(n, self.temb_dim, dtype=F32))
        lane_bias = ops.small_linear(emb, self.temb_w, self.temb_b, self._new(n, self.temb_w.shape[0], dtype=F32),
                                     silu_in=True)

        taps: Dict[str, torch.Tensor] = {}
        if taps_only and image_hidden_states is not None:
            raise ValueError("taps_only is the feature pass: no image_hidden_states")
        last_tap = self.last_tap_key if taps_only else None

        def attn(p, xa, hh, ww, key):
            image = None
            if img is not None:
                if key not in img:
                    raise KeyError(f"image_hidden_states is missing key {key!r}")
                image = img[key]
                if image.shape[2] != p["c"] or image.shape[1] % (hh * ww):
                    raise ValueError(f"image_hidden_states[{key!r}] has shape {tuple(image.shape)}; expected "
                                     f"(lanes, R*{hh * ww}, {p['c']})")
                if kv_map is None and image.shape[0] != n:
                    raise ValueError(f"image_hidden_states[{key!r}] has {image.shape[0]} lanes for a batch of {n}")
            out, tap = self._transformer(p, xa, text_kv, L, image, kv_map, 0, tap_only=(key == last_tap))
            if img is None:
                taps[key] = tap
            return out

        patches = ops.im2col3x3_nchw(sample, self._new(n * h * w, 64))
        x = self._new_act(n, h, w, boc0)
        # conv_in as a K = 64 GEMM over the pre-gathered patches, on the pixel grid (so that its GroupNorm partials
        # follow the grid's tiling)
        ops.gemm([Src.nhwc(patches.view(n, h, w, 64))], [(0, 0, 0, 0, 1)], (w, h, n), self.conv_in[0], x.t,
                 bias=self.conv_in[1], col_stats_out=x.stats())
        skips = [x]
        hh, ww = h, w
        for i, blk in enumerate(self.down):
            for j, rp in enumerate(blk["res"]):
                x = self._resnet(rp, x, None, lane_bias)
                if blk["attn"] is not None:
                    x = attn(blk["attn"][j], x, hh, ww, f"down_{i + 1}_{j + 1}")
                skips.append(x)
            if blk["down"] is not None:
                x = self._downsample(blk["down"], x)
                hh, ww = hh // 2, ww // 2
                skips.append(x)
        x = self._resnet(self.mid["res"][0], x, None, lane_bias)
        x = attn(self.mid["attn"][0], x, hh, ww, "mid")
        x = self._resnet(self.mid["res"][1], x, None, lane_bias)
        for i, blk in enumerate(self.up):
            for j, rp in enumerate(blk["res"]):
                x = self._resnet(rp, x, skips.pop(), lane_bias)
                if blk["attn"] is not None:
                    x = attn(blk["attn"][j], x, hh, ww, f"up_{i}_{j + 1}")
                    if x is None:  # taps_only: that was the last tap
                        return None, taps
            if blk["up"] is not None:
                x = self._upsample(blk["up"], x)
                hh, ww = hh * 2, ww * 2
        xn = self._gn(x, None, self.norm_out, self.eps, True)
        out = self._new(n, self.conv_out[2], hh, ww, dtype=F32)
        ops.gemm([Src.nhwc(xn.view(n, hh, ww, -1))], taps3x3(0, xn.shape[-1]), (ww, hh, n), self.conv_out[0], out,
                 bias=self.conv_out[1], mode=SG_EPI_NCHW_F32, out_channels=self.conv_out[2])
        return out, taps
