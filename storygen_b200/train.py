"""Training path of the UNet (train_StorySalon_stage2.py:258-332): stage 2 trains ONLY the image cross-attention
(`attn3`, :172-177), but its loss gradient has to travel back through everything downstream of the first attn3 layer.

What is hand-written here (this round): the attn3 attention core, forward AND backward, on the tcgen05 kernels
(`sg_attention` with its log-sum-exp rows, `sg_attention_bwd`: dQ / dK / dV with P recomputed tile by tile), wired in
as a `torch.autograd.Function`.  Everything else of the train-mode forward — the frozen convolutions, norms,
projections, the text / self attention cores — runs as stock torch ops under bf16 autocast so that torch autograd
provides their dgrad (and the wgrad of attn3's four projection matrices): scaffolding, as SURVEY.md §7 item 9 allows for
the first training round; the inference kernels are not differentiable.  Feature passes (no `image_hidden_states`)
involve no trainable parameter and still run on the inference engine.

Head dims 32 / 40 / 64 / 80 / 128 use the kernel; 160 (the two deepest SD-1.5 levels, < 2 % of the attention FLOPs) falls
to torch SDPA until the backward kernel's TMEM layout is extended.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import ops

BF16 = torch.bfloat16


class ImageAttentionFunction(torch.autograd.Function):
    """softmax(q k^T / sqrt(d)) v per head on the sm_100a kernels, differentiable."""

    @staticmethod
    def forward(ctx, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int):
        n, nq, c = q.shape
        nk = k.shape[1]
        hd = c // heads
        qb, kb, vb = q.to(BF16).contiguous(), k.to(BF16).contiguous(), v.to(BF16).contiguous()
        o = torch.empty_like(qb)
        lse = torch.empty(n, heads, nq, dtype=torch.float32, device=q.device)
        with torch.cuda.device(q.device):
            ops.attention(qb, kb, vb, o, n, heads, hd, nq, nk, (nq * c, c), (nk * c, c), (nk * c, c), lse=lse)
        ctx.save_for_backward(qb, kb, vb, o, lse)
        ctx.heads, ctx.dtypes = heads, (q.dtype, k.dtype, v.dtype)
        return o.to(q.dtype)

    @staticmethod
    def backward(ctx, d_o: torch.Tensor):
        qb, kb, vb, o, lse = ctx.saved_tensors
        with torch.cuda.device(qb.device):
            dq, dk, dv = ops.attention_bwd(qb, kb, vb, o, d_o.to(BF16).contiguous(), lse, ctx.heads)
        return dq.to(ctx.dtypes[0]), dk.to(ctx.dtypes[1]), dv.to(ctx.dtypes[2]), None


def _sdpa(q, k, v, heads):
    n, nq, c = q.shape
    hd = c // heads
    sp = lambda t: t.view(n, -1, heads, hd).transpose(1, 2)
    o = F.scaled_dot_product_attention(sp(q), sp(k), sp(v), scale=hd ** -0.5)
    return o.transpose(1, 2).reshape(n, nq, c)


def _attention(a, x, context=None, custom=False):
    """diffusers CrossAttention (model/attention.py:175-223): to_q / to_k / to_v, per-head softmax attention, to_out."""
    ctx = x if context is None else context
    q, k, v = a.to_q(x), a.to_k(ctx), a.to_v(ctx)
    hd = q.shape[-1] // a.heads
    if custom and hd in ops.ATTN_BWD_HEAD_DIMS:
        o = ImageAttentionFunction.apply(q, k, v, a.heads)
    else:
        o = _sdpa(q, k, v, a.heads)
    return a.to_out[0](o)


def _block(blk, h, text, image):
    """BasicTransformerBlock.forward (model/attention.py:236-302)."""
    h = _attention(blk.attn1, blk.norm1(h)) + h
    tap = h.clone()
    if image is not None:
        h_t = _attention(blk.attn2, blk.norm2(h), text) + h
        h_i = _attention(blk.attn3, blk.norm4(h), image, custom=True) + h
        h = h_t + h_i
    else:
        h = _attention(blk.attn2, blk.norm2(h), text) + h
    x = blk.norm3(h)
    proj = blk.ff.net[0].proj(x)
    val, gate = proj.chunk(2, dim=-1)
    return blk.ff.net[2](val * F.gelu(gate)) + h, tap


def _transformer(t, x, text, image):
    """Transformer2DModel.forward (model/attention.py:85-128) -> (output, tapped feature (n, hw, c))."""
    n, c, hh, ww = x.shape
    h = t.proj_in(t.norm(x))
    inner = h.shape[1]
    h = h.permute(0, 2, 3, 1).reshape(n, hh * ww, inner)
    h, tap = _block(t.transformer_blocks[0], h, text, image)
    h = h.reshape(n, hh, ww, inner).permute(0, 3, 1, 2)
    return t.proj_out(h) + x, tap


def _resnet(r, x, emb):
    h = r.conv1(F.silu(r.norm1(x)))
    h = h + r.time_emb_proj(F.silu(emb))[:, :, None, None]
    h = r.conv2(F.silu(r.norm2(h)))
    return (x if r.conv_shortcut is None else r.conv_shortcut(x)) + h


def _timestep_features(t, dim):
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    ang = t.float()[:, None] * freqs[None, :]
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1)


def unet_forward_train(m, sample, timestep, encoder_hidden_states, image_hidden_states: Optional[Dict[str, torch.Tensor]]):
    """UNet2DConditionModel.forward (model/unet_2d_condition.py:338-485) as differentiable torch ops + the custom attn3 core;
    returns (sample, taps)."""
    dev = sample.device
    n = sample.shape[0]
    t = timestep if torch.is_tensor(timestep) else torch.tensor([timestep], device=dev)
    t = t.to(dev).reshape(-1).expand(n)
    boc0 = m.config.block_out_channels[0]
    with torch.autocast("cuda", dtype=BF16):
        te = m.time_embedding
        emb = te.linear_2(F.silu(te.linear_1(_timestep_features(t, boc0).to(sample.dtype))))
        text, img = encoder_hidden_states, image_hidden_states
        taps: Dict[str, torch.Tensor] = {}

        def attn(tr, x, key):
            out, tap = _transformer(tr, x, text, None if img is None else img[key])
            if img is None:
                taps[key] = tap
            return out

        x = m.conv_in(sample)
        skips = [x]
        for i, blk in enumerate(m.down_blocks):
            for j, r in enumerate(blk.resnets):
                x = _resnet(r, x, emb)
                if blk.has_cross_attention:
                    x = attn(blk.attentions[j], x, f"down_{i + 1}_{j + 1}")
                skips.append(x)
            if blk.downsamplers is not None:
                x = blk.downsamplers[0].conv(x)
                skips.append(x)
        mb = m.mid_block
        x = _resnet(mb.resnets[0], x, emb)
        x = attn(mb.attentions[0], x, "mid")
        x = _resnet(mb.resnets[1], x, emb)
        for i, blk in enumerate(m.up_blocks):
            for j, r in enumerate(blk.resnets):
                x = _resnet(r, torch.cat([x, skips.pop()], dim=1), emb)
                if blk.has_cross_attention:
                    x = attn(blk.attentions[j], x, f"up_{i}_{j + 1}")
            if blk.upsamplers is not None:
                x = blk.upsamplers[0].conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))
        x = m.conv_out(F.silu(m.conv_norm_out(x)))
    return x.float(), taps
