"""DDIM scheduler with the interface the reference uses (model/pipeline.py:366-367, 420-424, 451, 461;
inference.py:48) and the arithmetic of diffusers==0.13.1 `DDIMScheduler` under the reference's
scheduler_config.json (scaled_linear betas 0.00085..0.012, 1000 train steps, steps_offset 1,
set_alpha_to_one False, clip_sample False, epsilon prediction).  Host-side scalar bookkeeping only:
the per-element work (add_noise, guidance + update) is done by the kernels in csrc/elementwise.cu."""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import List, Optional

import torch


@dataclass
class DDIMSchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None


class DDIMScheduler:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 beta_schedule: str = "scaled_linear", clip_sample: bool = False, set_alpha_to_one: bool = False,
                 steps_offset: int = 1, prediction_type: str = "epsilon", **unused):
        if beta_schedule != "scaled_linear" or prediction_type != "epsilon" or clip_sample:
            raise NotImplementedError("only the StoryGen scheduler configuration is implemented")
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0].clone()
        self.num_inference_steps: Optional[int] = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1, dtype=torch.long)

    @classmethod
    def from_pretrained(cls, path, subfolder: Optional[str] = None, **kw):
        p = os.path.join(path, subfolder) if subfolder else path
        with open(os.path.join(p, "scheduler_config.json")) as f:
            cfg = {k: v for k, v in json.load(f).items() if not k.startswith("_")}
        cfg.pop("skip_prk_steps", None)  # the shipped file is a PNDM config re-read as DDIM
        cfg.pop("trained_betas", None)
        cfg.update(kw)
        return cls(**cfg)

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        ts = [i * ratio + self.steps_offset for i in reversed(range(num_inference_steps))]
        self.timesteps = torch.tensor(ts, dtype=torch.long, device=device)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def step_alphas(self, timestep: int):
        """(alpha_bar_t, alpha_bar_prev) of the eta=0 update at `timestep`."""
        t_prev = int(timestep) - self.num_train_timesteps // self.num_inference_steps
        a_t = float(self.alphas_cumprod[int(timestep)])
        a_prev = float(self.alphas_cumprod[t_prev]) if t_prev >= 0 else float(self.final_alpha_cumprod)
        return a_t, a_prev

    def add_noise(self, original_samples, noise, timesteps):
        acp = self.alphas_cumprod.to(original_samples.device)
        t = torch.as_tensor(timesteps, device=original_samples.device).reshape(-1).long()
        a = acp[t].reshape(-1, *([1] * (original_samples.dim() - 1))).to(original_samples.dtype)
        return a.sqrt() * original_samples + (1 - a).sqrt() * noise

    def step(self, model_output, timestep, sample, eta: float = 0.0, **unused):
        if eta != 0.0:
            raise NotImplementedError("eta != 0 is not used by StoryGen")
        a_t, a_prev = self.step_alphas(int(timestep))
        x0 = (sample - (1 - a_t) ** 0.5 * model_output) / a_t ** 0.5
        return DDIMSchedulerOutput(prev_sample=a_prev ** 0.5 * x0 + (1 - a_prev) ** 0.5 * model_output,
                                   pred_original_sample=x0)
