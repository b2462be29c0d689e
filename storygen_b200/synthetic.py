"""Synthetic workload of the benchmark: per-image seeded latents and embeddings, keyed
by the GLOBAL image index so results do not depend on how images are sharded over ranks, plus a fast
on-device random initialisation of the UNet (no checkpoint exists in this environment)."""
from __future__ import annotations

import torch


def synthetic_inputs(batch: int, num_frames: int, latent_hw: int = 64, text_dim: int = 768, text_len: int = 77,
                     first_image: int = 0, pin: bool = False):
    def rn(shape, seed, scale=1.0):
        g = torch.Generator("cpu").manual_seed(seed)
        return torch.randn(shape, generator=g) * scale

    js = [first_image + j for j in range(batch)]
    out = dict(
        latents=torch.stack([rn((4, latent_hw, latent_hw), 1000 + j) for j in js]),
        text_cond=torch.stack([rn((text_len, text_dim), 2000 + j) for j in js]),
        text_uncond=rn((text_len, text_dim), 2999)[None].repeat(batch, 1, 1),
        zero_latents=rn((4, latent_hw, latent_hw), 3999, 0.1)[None].repeat(batch, 1, 1, 1),
        noise=torch.stack([rn((4, latent_hw, latent_hw), 4000 + j) for j in js]),
    )
    frames = torch.stack([rn((num_frames, 4, latent_hw, latent_hw), 3000 + j) for j in js])
    prev = torch.stack([rn((num_frames, text_len, text_dim), 5000 + j) for j in js])
    out["frame_latents"] = [frames[:, i].contiguous() for i in range(num_frames)]
    out["prev_cond"] = [prev[:, i].contiguous() for i in range(num_frames)]
    out["prev_uncond"] = [out["text_uncond"].clone() for _ in range(num_frames)]
    if pin:
        out = {k: ([t.pin_memory() for t in v] if isinstance(v, list) else v.pin_memory()) for k, v in out.items()}
    return out


def to_device(inputs, device):
    return {k: ([t.to(device) for t in v] if isinstance(v, list) else v.to(device)) for k, v in inputs.items()}


def input_bytes(inputs) -> int:
    n = 0
    for v in inputs.values():
        for t in (v if isinstance(v, list) else [v]):
            n += t.numel() * t.element_size()
    return n


def seeded_init_(model: torch.nn.Module, seed: int = 0, gain: float = 1.0):
    """Deterministic synthetic weights keyed by STATE-DICT NAME (CPU generator seeded by crc32 of the key), so any
    implementation of the 798-key UNet — this package, the test oracle, the reference — gets bit-identical tensors
    for the same seed regardless of construction order or device.  The benchmark uses it so that its output can be
    checked against a stored fp32 reference (tests/golden/bench_steps_fp32.pt)."""
    import zlib
    sd = model.state_dict()
    new = {}
    for key in sorted(sd):
        shape = tuple(sd[key].shape)
        g = torch.Generator("cpu").manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31))
        r = torch.randn(shape, generator=g, dtype=torch.float32)
        if len(shape) >= 2:  # conv / linear weight: variance preserving
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            w = r * (gain / fan_in ** 0.5)
        elif key.endswith("weight"):  # norm scale
            w = 1.0 + 0.1 * r
        else:  # bias / norm shift
            w = 0.05 * r
        new[key] = w
    model.load_state_dict(new)
    return model


def random_init_(model: torch.nn.Module, seed: int = 0):
    """Variance-preserving random weights, generated on the model's device (the values are irrelevant
    for throughput; parity tests use oracle/weights.py instead)."""
    g = torch.Generator(device=next(model.parameters()).device).manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() >= 2:
                fan_in = p[0].numel()
                p.normal_(0.0, fan_in ** -0.5, generator=g)
            elif name.endswith("weight"):
                p.normal_(1.0, 0.1, generator=g)
            else:
                p.normal_(0.0, 0.05, generator=g)
    return model
