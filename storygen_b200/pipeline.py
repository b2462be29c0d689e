"""StableDiffusionPipeline with the reference's interface (model/pipeline.py:22-491) whose denoising
loop — R feature passes + 1 denoise pass per step, dual-scale classifier-free guidance, eta=0 DDIM
update (model/pipeline.py:411-469) — runs on the sm_100a kernels.

Restructuring relative to the reference (results unchanged):
  * the R feature passes of a step are ONE UNet call over R*2B lanes with a per-lane timestep vector,
    lanes ordered (guidance lane, image, frame) so the 16 tapped feature maps land directly in the
    (lanes, R*HW, C) layout the image cross-attention reads — the reference's torch.cat over frames
    (model/pipeline.py:441-443) never happens;
  * the reference's third feature-pass CFG lane duplicates the second (model/pipeline.py:429-430);
    it is not recomputed: denoise-pass lanes [image] and [text+image] share one K/V lane (kv_map);
  * the feature pass stops at its last tapped layer: the reference keeps only `[1]` (the tapped maps) of that UNet
    call (model/pipeline.py:433-435) and drops the sample, so the rest of the last transformer block, conv_norm_out
    and conv_out of the feature pass are dead code (engine.forward(taps_only=True); ~1 % of a step's FLOPs);
  * a whole denoising step is captured once into a CUDA graph and replayed for the remaining steps.
Tokenizer / CLIP / VAE are the caller's torch modules and run either side of the hot path unchanged.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import ops
from .scheduler import DDIMScheduler

F32 = torch.float32
# SG_FEATURE_TAIL=1 runs the feature pass to its (unused) end, for A/B timing
_SKIP_FEATURE_TAIL = os.environ.get("SG_FEATURE_TAIL", "0") != "1"


@dataclass
class StableDiffusionPipelineOutput:
    images: object
    nsfw_content_detected: Optional[List[bool]] = None


def feature_timesteps(t: int, num_frames: int, stage: str) -> List[int]:
    """model/pipeline.py:414-427: ref_t = long(t / 10); 'auto-regressive' noises the oldest frame most."""
    ref_t = int(t / 10)
    if stage == "auto-regressive":
        return [ref_t * (num_frames - i) for i in range(num_frames)]
    if stage == "multi-image-condition":
        return [ref_t] * num_frames
    raise NotImplementedError(f"stage {stage!r}: only 'auto-regressive' and 'multi-image-condition' condition the "
                              "UNet on prior frames (model/pipeline.py:419-436)")


class _DenoiseLoop:
    """Static device buffers for one (B, R, latent shape, text shape, #steps, stage) configuration and one
    captured CUDA graph of a full denoising step; `load()` refills the buffers so the graph is reused
    across calls."""

    def __init__(self, unet, scheduler: DDIMScheduler, B, R, lat_shape, text_shape, num_inference_steps,
                 guidance_scale, image_guidance_scale, stage):
        dev = unet.device
        self.unet, self.eng = unet, unet.engine()
        self.B, self.R = B, R
        self.s_txt, self.s_img = float(guidance_scale), float(image_guidance_scale)
        scheduler.set_timesteps(num_inference_steps, device="cpu")
        self.timesteps = [int(t) for t in scheduler.timesteps]
        # feature-pass lanes: index ((g * B + b) * R + r), g = 0 "no image" lane (zero frame), g = 1 frame lane
        nf = 2 * B * R
        self.nf = nf
        self.latents = torch.empty(B, *lat_shape, dtype=F32, device=dev)
        self.f_src = torch.empty(2, B, R, *lat_shape, dtype=F32, device=dev)
        self.f_eps = torch.empty_like(self.f_src)
        self.f_text = torch.empty(2, B, R, *text_shape, dtype=torch.bfloat16, device=dev)
        self.f_noisy = torch.empty(nf, *lat_shape, dtype=F32, device=dev)
        self.m_text = torch.empty(3, B, *text_shape, dtype=torch.bfloat16, device=dev)
        self.lat3 = torch.empty(3 * B, *lat_shape, dtype=F32, device=dev)
        # attn2 K/V of every layer for the text lanes: projected once per load(), reused by all steps
        self.f_kv = torch.empty(nf * text_shape[0], self.eng.n_kv2, dtype=torch.bfloat16, device=dev)
        self.m_kv = torch.empty(3 * B * text_shape[0], self.eng.n_kv2, dtype=torch.bfloat16, device=dev)
        # denoise lane (g3, b): g3 = 0 -> K/V lane b (zero frame); g3 = 1, 2 -> K/V lane B + b
        self.kv_map = torch.tensor([b for b in range(B)] + [B + b for b in range(B)] * 2, dtype=torch.int32, device=dev)
        self.acp = scheduler.alphas_cumprod.to(dev, F32).contiguous()
        # per-step scalars, resident on the device: [f_t (nf) | m_t (3B) | a_t, a_prev] and int32 f_t
        S = len(self.timesteps)
        tab = torch.empty(S, nf + 3 * B + 2, dtype=F32)
        tab_i = torch.empty(S, nf, dtype=torch.int32)
        for s, t in enumerate(self.timesteps):
            ft = torch.tensor(feature_timesteps(t, R, stage), dtype=F32).repeat(2 * B)
            tab[s, :nf] = ft
            tab_i[s] = ft.to(torch.int32)
            tab[s, nf:nf + 3 * B] = float(t)
            a_t, a_prev = scheduler.step_alphas(t)
            tab[s, nf + 3 * B], tab[s, nf + 3 * B + 1] = a_t, a_prev
        self.tab, self.tab_i = tab.to(dev), tab_i.to(dev)
        self.par = torch.empty(nf + 3 * B + 2, dtype=F32, device=dev)
        self.par_i = torch.empty(nf, dtype=torch.int32, device=dev)
        # the reference indexes alphas_cumprod on the host and raises IndexError past the table (model/pipeline.py:420-424
        # via DDIMScheduler.add_noise); the add_noise kernel reads acp[t] unchecked, so the bound is enforced here
        if int(tab_i.max()) >= self.acp.numel() or int(tab_i.min()) < 0:
            raise IndexError(f"feature-pass timestep {int(tab_i.max())} is out of range for {self.acp.numel()} training "
                             f"timesteps (stage {stage!r} with {R} prior frames)")
        # GroupNorm statistics workspace of THIS loop: the captured graph bakes its address in, so it must not be
        # shared with (and re-allocated by) other callers of the same engine
        self.stats = ops.groupnorm_ws(max(nf, 3 * B), self.eng.groups, dev)
        self.graph = None
        self.kernels_per_step = 0
        self.kernel_launches = 0

    def load(self, latents, text_uncond, text_cond, prev_uncond, prev_cond, frame_latents, zero_latents, noise):
        """Copy one batch of inputs (host or device tensors) into the static buffers."""
        nb = dict(non_blocking=True)
        dev = self.latents.device

        def on_dev(t):
            # a host -> device copy_ that also changes the dtype converts on the CPU first (and then copies from pageable
            # memory, synchronously): move the bytes as they are (asynchronous from pinned memory), cast on the device
            return t if t.device == dev else t.to(dev, non_blocking=True)

        self.latents.copy_(latents, **nb)
        zero_latents, noise = on_dev(zero_latents), on_dev(noise)
        text_uncond, text_cond = on_dev(text_uncond), on_dev(text_cond)
        for r in range(self.R):
            self.f_src[0, :, r].copy_(zero_latents, **nb)
            self.f_src[1, :, r].copy_(frame_latents[r], **nb)
            self.f_eps[0, :, r].copy_(noise, **nb)
            self.f_eps[1, :, r].copy_(noise, **nb)
            self.f_text[0, :, r].copy_(on_dev(prev_uncond[r]), **nb)
            self.f_text[1, :, r].copy_(on_dev(prev_cond[r]), **nb)
        self.m_text[0].copy_(text_uncond, **nb)
        self.m_text[1].copy_(text_uncond, **nb)
        self.m_text[2].copy_(text_cond, **nb)
        with torch.cuda.device(self.latents.device):
            self._project()

    def _project(self):
        self.eng.project_text(self.f_text.view(self.nf, *self.f_text.shape[3:]), out=self.f_kv)
        self.eng.project_text(self.m_text.view(3 * self.B, *self.m_text.shape[2:]), out=self.m_kv)
        self.kernel_launches += 2

    def _step_body(self):
        nf, B = self.nf, self.B
        lat_shape = self.latents.shape[1:]
        ops.add_noise(self.f_src.view(nf, *lat_shape), self.f_eps.view(nf, *lat_shape), self.acp, self.par_i, self.f_noisy)
        # the feature pass only yields the 16 tapped maps (model/pipeline.py:431-436 takes `[1]`): its dead tail is not run
        taps = self.eng.forward(self.f_noisy, self.par[:nf], self.f_text.view(nf, *self.f_text.shape[3:]), None,
                                text_kv=self.f_kv, stats=self.stats, taps_only=_SKIP_FEATURE_TAIL)[1]
        cond = {k: v.view(2 * B, -1, v.shape[-1]) for k, v in taps.items()}
        self.lat3.view(3, B, -1).copy_(self.latents.view(1, B, -1).expand(3, B, -1))
        eps3 = self.eng.forward(self.lat3, self.par[nf:nf + 3 * B], self.m_text.view(3 * B, *self.m_text.shape[2:]), cond,
                                kv_map=self.kv_map, text_kv=self.m_kv, stats=self.stats)[0]
        ops.cfg_ddim_step(eps3, self.latents, self.s_img, self.s_txt, self.par[nf + 3 * B:])

    def run(self, use_cuda_graph: bool = True, callback=None, callback_steps: int = 1):
        with torch.cuda.device(self.latents.device):  # kernels launch on the CURRENT device's stream
            return self._run(use_cuda_graph, callback, callback_steps)

    def _run(self, use_cuda_graph, callback, callback_steps):
        from ._lib import launch_count
        for s, t in enumerate(self.timesteps):
            self.par.copy_(self.tab[s])
            self.par_i.copy_(self.tab_i[s])
            if self.graph is not None and use_cuda_graph:
                self.graph.replay()
                self.kernel_launches += self.kernels_per_step
            elif not use_cuda_graph or self.kernels_per_step == 0:
                # eager step (also the warm-up before capture: function attributes, allocator, tensor maps)
                c0 = launch_count()
                self._step_body()
                self.kernels_per_step = launch_count() - c0
                self.kernel_launches += self.kernels_per_step
            else:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._step_body()
                self.graph = g
                g.replay()
                self.kernel_launches += self.kernels_per_step
            if callback is not None and s % callback_steps == 0:
                callback(s, t, self.latents)
        return self.latents


class StableDiffusionPipeline:
    """Same constructor and __call__ as the reference's model/pipeline.py::StableDiffusionPipeline."""

    def __init__(self, vae, text_encoder, tokenizer, unet, scheduler):
        self.vae, self.text_encoder, self.tokenizer, self.unet, self.scheduler = vae, text_encoder, tokenizer, unet, scheduler
        boc = getattr(getattr(vae, "config", None), "block_out_channels", (128, 256, 512, 512))
        self.vae_scale_factor = 2 ** (len(boc) - 1)
        self._progress_bar_config: Dict = {}
        self.use_cuda_graph = True
        self._loops: Dict = {}
        self.kernel_launches = 0  # native kernels enqueued by denoise() so far (graph replays included)

    # ---- diffusers.DiffusionPipeline plumbing the reference's callers touch ----------------------
    @property
    def device(self):
        return self.unet.device

    _execution_device = device

    def set_progress_bar_config(self, **kwargs):
        self._progress_bar_config = kwargs

    def enable_xformers_memory_efficient_attention(self, *a, **k):
        return None  # the fused tcgen05 attention kernel is always on

    def enable_vae_slicing(self):
        return None

    @staticmethod
    def numpy_to_pil(images):
        from PIL import Image
        if images.ndim == 3:
            images = images[None, ...]
        images = (images * 255).round().astype("uint8")
        # model/pipeline.py:486-491: one LIST per image (the reference applies DiffusionPipeline.numpy_to_pil to each
        # image separately), which is why inference.py:117-120 saves `image[0]`
        return [[Image.fromarray(image)] for image in images]

    # ---- text / image conditioning (the steps either side of the hot path; caller's modules) -----
    def _encode_prompt(self, prompt, device, num_images_per_prompt, do_classifier_free_guidance, negative_prompt):
        """model/pipeline.py:87-196 -> (2B or B, L, D): [uncond; cond]."""
        batch_size = len(prompt) if isinstance(prompt, list) else 1
        tok = self.tokenizer(prompt, padding="max_length", max_length=self.tokenizer.model_max_length, truncation=True,
                             return_tensors="pt")
        emb = self.text_encoder(tok.input_ids.to(device), attention_mask=None)[0]
        bs, seq, _ = emb.shape
        emb = emb.repeat(1, num_images_per_prompt, 1).view(bs * num_images_per_prompt, seq, -1)
        if do_classifier_free_guidance:
            if negative_prompt is None:
                uncond_tokens = [""] * batch_size
            elif isinstance(negative_prompt, str):
                uncond_tokens = [negative_prompt]
            else:
                uncond_tokens = negative_prompt
            utok = self.tokenizer(uncond_tokens, padding="max_length", max_length=tok.input_ids.shape[-1],
                                  truncation=True, return_tensors="pt")
            un = self.text_encoder(utok.input_ids.to(device), attention_mask=None)[0]
            un = un.repeat(1, num_images_per_prompt, 1).view(batch_size * num_images_per_prompt, un.shape[1], -1)
            emb = torch.cat([un, emb])
        return emb

    def prepare_latents(self, batch_size, num_channels_latents, height, width, dtype, device, generator, latents=None):
        shape = (batch_size, num_channels_latents, height // self.vae_scale_factor, width // self.vae_scale_factor)
        if latents is None:
            if isinstance(generator, list):
                latents = torch.cat([torch.randn((1,) + shape[1:], generator=generator[i], device=device, dtype=dtype)
                                     for i in range(batch_size)], dim=0)
            else:
                latents = torch.randn(shape, generator=generator, device=device, dtype=dtype)
        elif latents.shape != shape:
            raise ValueError(f"Unexpected inputs shape, got {latents.shape}, expected {shape}")
        return latents.to(device) * self.scheduler.init_noise_sigma

    def decode_latents(self, latents):
        image = self.vae.decode((1 / 0.18215 * latents).to(next(self.vae.parameters()).dtype)).sample
        image = (image / 2 + 0.5).clamp(0, 1)
        return image.permute(0, 2, 3, 1).cpu().float().numpy()

    # ---- the hot path on synthetic / pre-encoded inputs ------------------------------------------
    @torch.no_grad()
    def denoise(self, latents, text_uncond, text_cond, prev_uncond: Sequence[torch.Tensor],
                prev_cond: Sequence[torch.Tensor], frame_latents: Sequence[torch.Tensor], zero_latents, noise,
                num_inference_steps: int = 50, guidance_scale: float = 7.0, image_guidance_scale: float = 3.5,
                stage: str = "auto-regressive", callback=None, callback_steps: int = 1) -> torch.Tensor:
        """Denoise B latents conditioned on R prior frames; every argument is a tensor (or list of R
        tensors) with leading dimension B.  Returns fp32 latents (B, 4, h, w) on the UNet's device."""
        if guidance_scale <= 1.0:
            raise NotImplementedError("guidance_scale <= 1 is broken in the reference (model/pipeline.py:430); "
                                      "only the classifier-free-guidance path is implemented")
        B, R = latents.shape[0], len(frame_latents)
        key = (B, R, tuple(latents.shape[1:]), tuple(text_cond.shape[1:]), num_inference_steps, float(guidance_scale),
               float(image_guidance_scale), stage, id(self.unet.engine()))
        loop = self._loops.get(key)
        if loop is None:
            self._loops.clear()  # one resident configuration: its graph pool holds every activation buffer
            loop = _DenoiseLoop(self.unet, self.scheduler, B, R, tuple(latents.shape[1:]), tuple(text_cond.shape[1:]),
                                num_inference_steps, guidance_scale, image_guidance_scale, stage)
            self._loops[key] = loop
        loop.load(latents, text_uncond, text_cond, prev_uncond, prev_cond, frame_latents, zero_latents, noise)
        out = loop.run(self.use_cuda_graph, callback, callback_steps)
        self.kernel_launches = loop.kernel_launches
        return out.clone()

    # ---- reference-compatible entry point ---------------------------------------------------------
    @torch.no_grad()
    def __call__(self, stage, prompt: Union[str, List[str]], image_prompt=None, prev_prompt=None,
                 height: Optional[int] = None, width: Optional[int] = None, num_inference_steps: int = 50,
                 guidance_scale: float = 7.5, image_guidance_scale: float = 3.5, negative_prompt=None,
                 num_images_per_prompt: Optional[int] = 1, eta: float = 0.0, generator=None, latents=None,
                 output_type: Optional[str] = "pil", return_dict: bool = True, callback: Optional[Callable] = None,
                 callback_steps: Optional[int] = 1):
        """model/pipeline.py:273-484.  The global-RNG draws happen in the reference's order: VAE posterior
        of the zero image, of each prior frame, then one `noise` tensor shared by all frames and steps."""
        height = height or self.unet.config.sample_size * self.vae_scale_factor
        width = width or self.unet.config.sample_size * self.vae_scale_factor
        if height % 8 or width % 8:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")
        if eta != 0.0:
            raise NotImplementedError("eta != 0")
        batch_size = 1 if isinstance(prompt, str) else len(prompt)
        device = self._execution_device
        nipp = num_images_per_prompt
        te = self._encode_prompt(prompt, device, nipp, True, negative_prompt)
        prev = [self._encode_prompt(p, device, nipp, True, negative_prompt) for p in prev_prompt]
        B = batch_size * nipp
        latents = self.prepare_latents(B, self.unet.in_channels, height, width, te.dtype, device, generator, latents)
        vae_dtype = next(self.vae.parameters()).dtype
        image_prompt = image_prompt.to(device=device, dtype=vae_dtype)
        frames = torch.transpose(image_prompt, 0, 1)
        zero = self.vae.encode(frames[0] * 0).latent_dist.sample() * 0.18215
        zero = zero.repeat(nipp, 1, 1, 1)
        frame_latents = [(self.vae.encode(fr).latent_dist.sample() * 0.18215).repeat(nipp, 1, 1, 1) for fr in frames]
        noise = torch.randn_like(frame_latents[0])
        out = self.denoise(latents, te[:B], te[B:], [p[:B] for p in prev], [p[B:] for p in prev], frame_latents, zero,
                           noise, num_inference_steps, guidance_scale, image_guidance_scale, stage, callback,
                           callback_steps or 1)
        image = self.decode_latents(out)
        if output_type == "pil":
            image = self.numpy_to_pil(image)
        if not return_dict:
            return (image, None)
        return StableDiffusionPipelineOutput(images=image, nsfw_content_detected=None)
