"""ORACLE (test infrastructure, not product code).

fp32 CPU restatement of StoryGen's denoising UNet, written from SURVEY.md / the reference's behaviour,
with the reference's exact state-dict keys so weights can be exchanged with both the reference
(tests/test_oracle_vs_reference.py, via oracle/ref_shim) and the CUDA path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this package.  The product package `storygen_b200` never does.

Parity status: the reference ships no tests or golden vectors.  This restatement is
pinned against the reference's own `model/*.py`, imported unmodified through `oracle/ref_shim`
(a test-only stand-in for the absent `diffusers==0.13.1`), by tests/test_oracle_vs_reference.py and
by the fixtures under tests/golden/ that `oracle/make_golden.py` generated from the REFERENCE code.

Reference call sites restated here (paths relative to /root/reference):
  model/unet_2d_condition.py:118-270 (construction), :338-485 (forward), :487-510 (load_SDM_state_dict)
  model/unet_2d_blocks.py:197-297 (mid), :300-436 (CrossAttnDownBlock2D), :439-515 (DownBlock2D),
                          :518-660 (CrossAttnUpBlock2D), :663-733 (UpBlock2D)
  model/attention.py:26-128 (Transformer2DModel), :131-302 (BasicTransformerBlock), :305-393 (FF/GEGLU)
  diffusers==0.13.1 (not vendored): CrossAttention, ResnetBlock2D,
  Downsample2D, Upsample2D, Timesteps, TimestepEmbedding.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

SD15_CONFIG = dict(
    in_channels=4,
    out_channels=4,
    block_out_channels=(320, 640, 1280, 1280),
    down_block_types=("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
    up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
    layers_per_block=2,
    norm_num_groups=32,
    norm_eps=1e-5,
    cross_attention_dim=768,
    attention_head_dim=8,  # diffusers quirk: this is the NUMBER of heads
)


def timestep_features(t: torch.Tensor, dim: int) -> torch.Tensor:
    """diffusers Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0): [cos | sin], fp32."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    ang = t.float()[:, None] * freqs[None, :]
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_dim: int, dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class CrossAttention(nn.Module):
    """diffusers 0.13.1 CrossAttention, default processor (scores -> softmax -> weighted sum)."""

    def __init__(self, query_dim: int, cross_attention_dim: Optional[int], heads: int, dim_head: int):
        super().__init__()
        inner = heads * dim_head
        kv_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads, self.dim_head = heads, dim_head
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(kv_dim, inner, bias=False)
        self.to_v = nn.Linear(kv_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])

    def forward(self, x, context=None):
        ctx = x if context is None else context
        b, nq, _ = x.shape
        h, d = self.heads, self.dim_head
        q = self.to_q(x).view(b, nq, h, d).transpose(1, 2)
        k = self.to_k(ctx).view(b, ctx.shape[1], h, d).transpose(1, 2)
        v = self.to_v(ctx).view(b, ctx.shape[1], h, d).transpose(1, 2)
        # mathematically identical to baddbmm(alpha=d^-0.5) -> softmax -> bmm, without materialising
        # the (b*h, nq, nk) score tensor (4.8 GB for one L0 attn3 call on CPU)
        o = F.scaled_dot_product_attention(q, k, v, scale=d ** -0.5)
        o = o.transpose(1, 2).reshape(b, nq, h * d)
        return self.to_out[0](o)


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        val, gate = self.proj(x).chunk(2, dim=-1)
        return val * F.gelu(gate)  # erf GELU


class FeedForward(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Dropout(0.0), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class BasicTransformerBlock(nn.Module):
    """model/attention.py:236-302 — self-attn, tap, text + image cross-attn (residual counted twice), FF."""

    def __init__(self, dim: int, heads: int, dim_head: int, cross_attention_dim: int):
        super().__init__()
        self.attn1 = CrossAttention(dim, None, heads, dim_head)
        self.norm1 = nn.LayerNorm(dim)
        self.attn2 = CrossAttention(dim, cross_attention_dim, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim)
        self.attn3 = CrossAttention(dim, dim, heads, dim_head)
        self.norm4 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)
        self.norm3 = nn.LayerNorm(dim)

    def forward(self, h, text, image=None):
        h = self.attn1(self.norm1(h)) + h
        tap = h.clone()
        h_t = self.attn2(self.norm2(h), text) + h
        if image is not None:
            h_i = self.attn3(self.norm4(h), image) + h  # keys/values are the RAW features
            h = h_t + h_i  # = 2h + a_t + a_i
        else:
            h = h_t
        h = self.ff(self.norm3(h)) + h
        return h, tap


class Transformer2DModel(nn.Module):
    def __init__(self, heads: int, dim_head: int, in_channels: int, cross_attention_dim: int, groups: int):
        super().__init__()
        inner = heads * dim_head
        self.norm = nn.GroupNorm(groups, in_channels, eps=1e-6)
        self.proj_in = nn.Conv2d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(inner, heads, dim_head, cross_attention_dim)])
        self.proj_out = nn.Conv2d(inner, in_channels, 1)

    def forward(self, x, text, image=None):
        b, c, hh, ww = x.shape
        h = self.proj_in(self.norm(x))
        h = h.permute(0, 2, 3, 1).reshape(b, hh * ww, -1)
        h, tap = self.transformer_blocks[0](h, text, image)
        h = h.reshape(b, hh, ww, -1).permute(0, 3, 1, 2)
        return self.proj_out(h) + x, tap


class ResnetBlock2D(nn.Module):
    def __init__(self, cin: int, cout: int, temb: int, groups: int, eps: float):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, emb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(emb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Downsample2D(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class _Block(nn.Module):
    """resnets (+ attentions) (+ one down/up-sampler); registered names match the reference."""

    def __init__(self):
        super().__init__()


def _make_down(cin, cout, temb, groups, eps, heads, xdim, layers, attn: bool, down: bool):
    blk = _Block()
    blk.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, temb, groups, eps) for i in range(layers)])
    if attn:
        blk.attentions = nn.ModuleList([Transformer2DModel(heads, cout // heads, cout, xdim, groups) for _ in range(layers)])
    if down:
        blk.downsamplers = nn.ModuleList([Downsample2D(cout)])
    blk.has_attn, blk.has_down = attn, down
    return blk


def _make_up(cin, cout, prev, temb, groups, eps, heads, xdim, layers, attn: bool, up: bool):
    blk = _Block()
    res = []
    for i in range(layers):
        skip = cin if i == layers - 1 else cout
        rin = prev if i == 0 else cout
        res.append(ResnetBlock2D(rin + skip, cout, temb, groups, eps))
    blk.resnets = nn.ModuleList(res)
    if attn:
        blk.attentions = nn.ModuleList([Transformer2DModel(heads, cout // heads, cout, xdim, groups) for _ in range(layers)])
    if up:
        blk.upsamplers = nn.ModuleList([Upsample2D(cout)])
    blk.has_attn, blk.has_up = attn, up
    return blk


class OracleUNet(nn.Module):
    """UNet2DConditionModel restated.  forward(sample, timestep, text, image_hidden_states=None) ->
    (sample, dict).  Feature pass (image_hidden_states is None) returns the 16 taps keyed
    down_{1..3}_{1,2}, mid, up_{1..3}_{1..3}; denoise pass returns an empty dict.

    Cond keys are chosen by BLOCK INDEX.  The reference derives them from the latent height
    (model/unet_2d_blocks.py:380-381, 600-601), which coincides with the block index only for latent
    heights 64..79 and crashes elsewhere."""

    def __init__(self, **cfg):
        super().__init__()
        c = dict(SD15_CONFIG)
        c.update(cfg)
        self.cfg = c
        boc = c["block_out_channels"]
        groups, eps, heads, xdim = c["norm_num_groups"], c["norm_eps"], c["attention_head_dim"], c["cross_attention_dim"]
        temb = boc[0] * 4
        self.conv_in = nn.Conv2d(c["in_channels"], boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], temb)
        self.down_blocks = nn.ModuleList()
        out = boc[0]
        for i, t in enumerate(c["down_block_types"]):
            cin, out = out, boc[i]
            self.down_blocks.append(_make_down(cin, out, temb, groups, eps, heads, xdim, c["layers_per_block"],
                                               t == "CrossAttnDownBlock2D", i != len(boc) - 1))
        mid = _Block()
        mid.resnets = nn.ModuleList([ResnetBlock2D(boc[-1], boc[-1], temb, groups, eps) for _ in range(2)])
        mid.attentions = nn.ModuleList([Transformer2DModel(heads, boc[-1] // heads, boc[-1], xdim, groups)])
        self.mid_block = mid
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(boc))
        out = rev[0]
        for i, t in enumerate(c["up_block_types"]):
            prev, out = out, rev[i]
            cin = rev[min(i + 1, len(boc) - 1)]
            self.up_blocks.append(_make_up(cin, out, prev, temb, groups, eps, heads, xdim, c["layers_per_block"] + 1,
                                           t == "CrossAttnUpBlock2D", i != len(boc) - 1))
        self.conv_norm_out = nn.GroupNorm(groups, boc[0], eps=eps)
        self.conv_out = nn.Conv2d(boc[0], c["out_channels"], 3, padding=1)

    @property
    def in_channels(self):
        return self.cfg["in_channels"]

    def forward(self, sample, timestep, encoder_hidden_states, image_hidden_states: Optional[Dict[str, torch.Tensor]] = None):
        n = sample.shape[0]
        t = timestep if torch.is_tensor(timestep) else torch.tensor([timestep], device=sample.device)
        if t.dim() == 0:
            t = t[None]
        t = t.to(sample.device).expand(n)
        boc = self.cfg["block_out_channels"]
        emb = self.time_embedding(timestep_features(t, boc[0]).to(sample.dtype))
        text, img = encoder_hidden_states, image_hidden_states
        taps: Dict[str, torch.Tensor] = {}
        h = self.conv_in(sample)
        skips = [h]
        for i, blk in enumerate(self.down_blocks):
            for j, res in enumerate(blk.resnets):
                h = res(h, emb)
                if blk.has_attn:
                    key = f"down_{i + 1}_{j + 1}"
                    h, tap = blk.attentions[j](h, text, None if img is None else img[key])
                    if img is None:
                        taps[key] = tap
                skips.append(h)
            if blk.has_down:
                h = blk.downsamplers[0](h)
                skips.append(h)
        h = self.mid_block.resnets[0](h, emb)
        h, tap = self.mid_block.attentions[0](h, text, None if img is None else img["mid"])
        if img is None:
            taps["mid"] = tap
        h = self.mid_block.resnets[1](h, emb)
        for i, blk in enumerate(self.up_blocks):
            for j, res in enumerate(blk.resnets):
                h = res(torch.cat([h, skips.pop()], dim=1), emb)
                if blk.has_attn:
                    key = f"up_{i}_{j + 1}"
                    h, tap = blk.attentions[j](h, text, None if img is None else img[key])
                    if img is None:
                        taps[key] = tap
            if blk.has_up:
                h = blk.upsamplers[0](h)
        h = self.conv_out(F.silu(self.conv_norm_out(h)))
        return h, taps

    def load_SDM_state_dict(self, sd):
        """model/unet_2d_condition.py:487-510 — missing attn3/norm4 keys are filled from attn1/norm1."""
        own = self.state_dict()
        sd = {k: v for k, v in sd.items() if k in own and v.shape == own[k].shape}
        for k in own:
            if k not in sd:
                sd[k] = sd[k.replace("attn3", "attn1").replace("norm4", "norm1")]
        self.load_state_dict(sd)


# small SD-shaped config used by the fast tests: head dims 32 / 64 / 128 (2 heads per level)
TINY_CONFIG = dict(block_out_channels=(64, 128, 256, 256), norm_num_groups=8, cross_attention_dim=64,
                   attention_head_dim=2)
