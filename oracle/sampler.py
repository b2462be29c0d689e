"""ORACLE (test infrastructure): DDIM scheduler arithmetic and the StoryGen denoising loop, fp32.

Restates diffusers==0.13.1 `DDIMScheduler` (absent here) as configured by
/root/reference/ckpt/stable-diffusion-v1-5/scheduler/scheduler_config.json:4-12, and the loop of
/root/reference/model/pipeline.py:411-469 (R feature passes + 1 denoise pass per step, dual-scale
classifier-free guidance, eta=0 update) on SYNTHETIC latents / embeddings, i.e. without tokenizer,
CLIP, VAE or PIL.  Checked against the reference's own `StableDiffusionPipeline.__call__` (run through
oracle/ref_shim with stub tokenizer / text encoder / VAE) by tests/test_oracle_vs_reference.py.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch

SCHEDULER_CONFIG = dict(beta_start=0.00085, beta_end=0.012, num_train_timesteps=1000, steps_offset=1,
                        set_alpha_to_one=False, clip_sample=False)


def alphas_cumprod(cfg=SCHEDULER_CONFIG) -> torch.Tensor:
    betas = torch.linspace(cfg["beta_start"] ** 0.5, cfg["beta_end"] ** 0.5, cfg["num_train_timesteps"],
                           dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def ddim_timesteps(num_inference_steps: int, cfg=SCHEDULER_CONFIG) -> List[int]:
    ratio = cfg["num_train_timesteps"] // num_inference_steps
    return [i * ratio + cfg["steps_offset"] for i in reversed(range(num_inference_steps))]


def add_noise(x: torch.Tensor, eps: torch.Tensor, t: torch.Tensor, acp: torch.Tensor) -> torch.Tensor:
    a = acp[t].reshape(-1, *([1] * (x.dim() - 1)))
    return a.sqrt() * x + (1 - a).sqrt() * eps


def ddim_step(eps: torch.Tensor, t: int, x: torch.Tensor, num_inference_steps: int, acp: torch.Tensor,
              cfg=SCHEDULER_CONFIG) -> torch.Tensor:
    t_prev = t - cfg["num_train_timesteps"] // num_inference_steps
    a_t = acp[t]
    a_prev = acp[t_prev] if t_prev >= 0 else acp[0]  # set_alpha_to_one=False -> final alpha = acp[0]
    x0 = (x - (1 - a_t).sqrt() * eps) / a_t.sqrt()
    return a_prev.sqrt() * x0 + (1 - a_prev).sqrt() * eps


def feature_timesteps(t: int, num_frames: int, stage: str) -> List[int]:
    """model/pipeline.py:414-427 — ref_t = long(t / 10); auto-regressive: the oldest frame is noised most."""
    ref_t = int(t / 10)
    if stage == "auto-regressive":
        return [ref_t * (num_frames - i) for i in range(num_frames)]
    if stage == "multi-image-condition":
        return [ref_t] * num_frames
    raise ValueError(f"unsupported stage {stage!r}")


@torch.no_grad()
def denoise(unet: Callable, latents: torch.Tensor, text_uncond: torch.Tensor, text_cond: torch.Tensor,
            prev_uncond: Sequence[torch.Tensor], prev_cond: Sequence[torch.Tensor],
            frame_latents: Sequence[torch.Tensor], zero_latents: torch.Tensor, noise: torch.Tensor,
            num_inference_steps: int = 50, guidance_scale: float = 7.0, image_guidance_scale: float = 3.5,
            stage: str = "auto-regressive", callback: Optional[Callable] = None) -> torch.Tensor:
    """The loop of model/pipeline.py:411-469 for B images with R prior frames.

    latents (B,4,h,w); text_* (B,77,D); prev_*[i] (B,77,D) text of prior frame i; frame_latents[i],
    zero_latents, noise (B,4,h,w).  `unet(sample, timestep, text, image_hidden_states)` returns
    (sample, taps).  CFG lanes: [no text, no image | no text, image | text, image]."""
    acp = alphas_cumprod().to(latents.device)
    R = len(frame_latents)
    B = latents.shape[0]
    for t in ddim_timesteps(num_inference_steps):
        conds: List[Dict[str, torch.Tensor]] = []
        for i, ft in enumerate(feature_timesteps(t, R, stage)):
            tt = torch.full((B,), ft, dtype=torch.long, device=latents.device)
            noisy_frame = add_noise(frame_latents[i], noise, tt, acp)
            noisy_zero = add_noise(zero_latents, noise, tt, acp)
            x = torch.cat([noisy_zero, noisy_frame, noisy_frame])
            txt = torch.cat([prev_uncond[i], prev_cond[i], prev_cond[i]])
            conds.append(unet(x, ft, txt, None)[1])
        image_hidden_states = {k: torch.cat([c[k] for c in conds], dim=1) for k in conds[0]}
        txt = torch.cat([text_uncond, text_uncond, text_cond])
        eps = unet(torch.cat([latents] * 3), t, txt, image_hidden_states)[0]
        e_u, e_i, e_a = eps.chunk(3)
        eps = e_u + image_guidance_scale * (e_i - e_u) + guidance_scale * (e_a - e_i)
        latents = ddim_step(eps, t, latents, num_inference_steps, acp)
        if callback is not None:
            callback(t, latents)
    return latents


def synthetic_inputs(batch: int, num_frames: int, latent_hw: int = 64, text_dim: int = 768, text_len: int = 77,
                     device="cpu", first_image: int = 0):
    """The synthetic workload of SURVEY.md §8(d): per-image seeded latents / embeddings, so results
    do not depend on how images are sharded over ranks."""
    def rn(shape, seed, scale=1.0):
        g = torch.Generator("cpu").manual_seed(seed)
        return (torch.randn(shape, generator=g) * scale).to(device)

    js = [first_image + j for j in range(batch)]
    lat = torch.stack([rn((4, latent_hw, latent_hw), 1000 + j) for j in js])
    text_cond = torch.stack([rn((text_len, text_dim), 2000 + j) for j in js])
    text_uncond = rn((text_len, text_dim), 2999)[None].repeat(batch, 1, 1)
    frames_all = torch.stack([rn((num_frames, 4, latent_hw, latent_hw), 3000 + j) for j in js])  # (B,R,...)
    zero = rn((4, latent_hw, latent_hw), 3999, 0.1)[None].repeat(batch, 1, 1, 1)
    noise = torch.stack([rn((4, latent_hw, latent_hw), 4000 + j) for j in js])
    prev_all = torch.stack([rn((num_frames, text_len, text_dim), 5000 + j) for j in js])  # (B,R,77,D)
    return dict(latents=lat, text_uncond=text_uncond, text_cond=text_cond,
                prev_uncond=[text_uncond.clone() for _ in range(num_frames)],
                prev_cond=[prev_all[:, i].contiguous() for i in range(num_frames)],
                frame_latents=[frames_all[:, i].contiguous() for i in range(num_frames)],
                zero_latents=zero, noise=noise)
