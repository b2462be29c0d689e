import torch
import torch.nn as nn


class CrossAttention(nn.Module):
    """diffusers 0.13.1 models/cross_attention.py::CrossAttention with the default CrossAttnProcessor:
    q/k/v bias-free Linear, head split, baddbmm(alpha=scale) -> softmax -> bmm, to_out = [Linear, Dropout]."""

    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0, bias=False,
                 upcast_attention=False, upcast_softmax=False, added_kv_proj_dim=None, norm_num_groups=None,
                 processor=None):
        super().__init__()
        inner_dim = dim_head * heads
        cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner_dim, bias=bias)
        self.to_k = nn.Linear(cross_attention_dim, inner_dim, bias=bias)
        self.to_v = nn.Linear(cross_attention_dim, inner_dim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner_dim, query_dim), nn.Dropout(dropout)])

    def head_to_batch_dim(self, t):
        b, n, c = t.shape
        h = self.heads
        return t.reshape(b, n, h, c // h).permute(0, 2, 1, 3).reshape(b * h, n, c // h)

    def batch_to_head_dim(self, t):
        bh, n, d = t.shape
        h = self.heads
        return t.reshape(bh // h, h, n, d).permute(0, 2, 1, 3).reshape(bh // h, n, d * h)

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        q = self.to_q(hidden_states)
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        k = self.to_k(ctx)
        v = self.to_v(ctx)
        q, k, v = self.head_to_batch_dim(q), self.head_to_batch_dim(k), self.head_to_batch_dim(v)
        scores = torch.baddbmm(
            torch.empty(q.shape[0], q.shape[1], k.shape[1], dtype=q.dtype, device=q.device),
            q, k.transpose(-1, -2), beta=0, alpha=self.scale)
        probs = scores.softmax(dim=-1).to(q.dtype)
        out = self.batch_to_head_dim(torch.bmm(probs, v))
        out = self.to_out[0](out)
        return self.to_out[1](out)
