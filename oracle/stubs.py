"""ORACLE (test infrastructure): deterministic stand-ins for tokenizer / CLIP text encoder / VAE so
the reference's StableDiffusionPipeline.__call__ (model/pipeline.py:273-484) can run end to end
without the weights that are absent here.  Shared by oracle/make_golden.py and the tests."""
from __future__ import annotations

import zlib
from types import SimpleNamespace

import torch
import torch.nn as nn


class StubTokenizer:
    model_max_length = 77

    def __call__(self, prompt, padding=None, max_length=None, truncation=None, return_tensors=None):
        prompts = [prompt] if isinstance(prompt, str) else list(prompt)
        L = max_length or self.model_max_length
        ids = torch.zeros(len(prompts), L, dtype=torch.long)
        for i, p in enumerate(prompts):
            for j, w in enumerate(p.split()[:L]):
                ids[i, j] = 1 + zlib.crc32(w.encode()) % 4095
        return SimpleNamespace(input_ids=ids)

    def batch_decode(self, ids):
        return ["" for _ in ids]


class StubTextEncoder(nn.Module):
    def __init__(self, dim: int, seed: int = 7):
        super().__init__()
        g = torch.Generator("cpu").manual_seed(seed)
        self.table = nn.Parameter(torch.randn(4096, dim, generator=g), requires_grad=False)
        self.pos = nn.Parameter(0.3 * torch.randn(77, dim, generator=g), requires_grad=False)
        self.config = SimpleNamespace()

    def forward(self, input_ids, attention_mask=None):
        return (self.table[input_ids] + self.pos[None, : input_ids.shape[1]],)


class StubVAE(nn.Module):
    """encode: 8x average pool of the image -> 4 channels, plus posterior noise from the GLOBAL RNG
    (as AutoencoderKL's latent_dist.sample() without a generator does); decode: nearest 8x upsample."""

    def __init__(self):
        super().__init__()
        self.mix = nn.Parameter(torch.tensor([[0.9, 0.1, 0.0], [0.0, 0.9, 0.1], [0.1, 0.0, 0.9], [0.3, 0.3, 0.3]]),
                                requires_grad=False)
        self.config = SimpleNamespace(block_out_channels=(128, 256, 512, 512))

    def encode(self, x):
        mean = torch.einsum("oc,bchw->bohw", self.mix.to(x.dtype), torch.nn.functional.avg_pool2d(x, 8)) * 4 - 2
        return SimpleNamespace(latent_dist=SimpleNamespace(sample=lambda: mean + 0.2 * torch.randn_like(mean)))

    def decode(self, z):
        img = torch.nn.functional.interpolate(z[:, :3], scale_factor=8.0, mode="nearest")
        return SimpleNamespace(sample=img)
