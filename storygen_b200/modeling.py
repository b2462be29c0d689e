"""Module tree of StoryGen's UNet with the reference's names, constructor and state-dict schema.

These classes are PARAMETER CONTAINERS: they own the 798 tensors under exactly the keys of the
reference model (model/unet_2d_condition.py:118-270, model/unet_2d_blocks.py, model/attention.py and
the diffusers 0.13.1 layers they instantiate), so checkpoints load unchanged and
`named_modules()` / `parameters()` behave as the training scripts expect
(train_StorySalon_stage2.py:168-177).  The arithmetic lives in `engine.UNetEngine`, which runs the
hand-written sm_100a kernels; there is no torch forward here.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import Any, Dict, Optional, Tuple, Union

import torch
import torch.nn as nn


class _Config(dict):
    """dict with attribute access (diffusers FrozenDict behaviour that callers rely on)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels: int, time_embed_dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)


class CrossAttention(nn.Module):
    """Parameters of diffusers' CrossAttention as built at model/attention.py:175-223."""

    def __init__(self, query_dim: int, cross_attention_dim: Optional[int] = None, heads: int = 8, dim_head: int = 64,
                 bias: bool = False):
        super().__init__()
        inner = heads * dim_head
        kv = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads, self.dim_head = heads, dim_head
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(kv, inner, bias=bias)
        self.to_v = nn.Linear(kv, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class FeedForward(nn.Module):
    def __init__(self, dim: int, mult: int = 4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])


class BasicTransformerBlock(nn.Module):
    """attn1 (self) / attn2 (text) / attn3 (prior-frame image) + norms 1-4 + GEGLU FF
    (model/attention.py:131-234)."""

    def __init__(self, dim: int, num_attention_heads: int, attention_head_dim: int, cross_attention_dim: int):
        super().__init__()
        self.attn1 = CrossAttention(dim, None, num_attention_heads, attention_head_dim)
        self.norm1 = nn.LayerNorm(dim)
        self.attn2 = CrossAttention(dim, cross_attention_dim, num_attention_heads, attention_head_dim)
        self.norm2 = nn.LayerNorm(dim)
        self.attn3 = CrossAttention(dim, dim, num_attention_heads, attention_head_dim)
        self.norm4 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)
        self.norm3 = nn.LayerNorm(dim)


class Transformer2DModel(nn.Module):
    def __init__(self, num_attention_heads: int, attention_head_dim: int, in_channels: int,
                 cross_attention_dim: int, norm_num_groups: int = 32):
        super().__init__()
        inner = num_attention_heads * attention_head_dim
        self.num_attention_heads, self.attention_head_dim = num_attention_heads, attention_head_dim
        self.norm = nn.GroupNorm(norm_num_groups, in_channels, eps=1e-6, affine=True)  # model/attention.py:55
        self.proj_in = nn.Conv2d(in_channels, inner, kernel_size=1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, num_attention_heads, attention_head_dim, cross_attention_dim)])
        self.proj_out = nn.Conv2d(inner, in_channels, kernel_size=1)


class ResnetBlock2D(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, temb_channels: int, groups: int, eps: float):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None


class Downsample2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=1)


class Upsample2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)


def _transformers(n, heads, channels, xdim, groups):
    return nn.ModuleList([Transformer2DModel(heads, channels // heads, channels, xdim, groups) for _ in range(n)])


class CrossAttnDownBlock2D(nn.Module):
    has_cross_attention = True

    def __init__(self, in_channels, out_channels, temb_channels, num_layers, eps, groups, heads, xdim, add_downsample):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, temb_channels, groups, eps)
             for i in range(num_layers)])
        self.attentions = _transformers(num_layers, heads, out_channels, xdim, groups)
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels)]) if add_downsample else None


class DownBlock2D(nn.Module):
    has_cross_attention = False

    def __init__(self, in_channels, out_channels, temb_channels, num_layers, eps, groups, add_downsample):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, temb_channels, groups, eps)
             for i in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels)]) if add_downsample else None


class UNetMidBlock2DCrossAttn(nn.Module):
    has_cross_attention = True

    def __init__(self, in_channels, temb_channels, eps, groups, heads, xdim):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels, in_channels, temb_channels, groups, eps) for _ in range(2)])
        self.attentions = _transformers(1, heads, in_channels, xdim, groups)


def _up_resnets(in_channels, out_channels, prev_output_channel, temb_channels, num_layers, groups, eps):
    res = []
    for i in range(num_layers):
        skip = in_channels if i == num_layers - 1 else out_channels
        rin = prev_output_channel if i == 0 else out_channels
        res.append(ResnetBlock2D(rin + skip, out_channels, temb_channels, groups, eps))
    return nn.ModuleList(res)


class CrossAttnUpBlock2D(nn.Module):
    has_cross_attention = True

    def __init__(self, in_channels, out_channels, prev_output_channel, temb_channels, num_layers, eps, groups, heads,
                 xdim, add_upsample):
        super().__init__()
        self.resnets = _up_resnets(in_channels, out_channels, prev_output_channel, temb_channels, num_layers, groups, eps)
        self.attentions = _transformers(num_layers, heads, out_channels, xdim, groups)
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels)]) if add_upsample else None


class UpBlock2D(nn.Module):
    has_cross_attention = False

    def __init__(self, in_channels, out_channels, prev_output_channel, temb_channels, num_layers, eps, groups,
                 add_upsample):
        super().__init__()
        self.resnets = _up_resnets(in_channels, out_channels, prev_output_channel, temb_channels, num_layers, groups, eps)
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels)]) if add_upsample else None


@dataclass
class UNet2DConditionOutput:
    """model/unet_2d_condition.py:24-32"""
    sample: torch.Tensor
    img_dif_conditions: Dict[str, torch.Tensor]

    def __getitem__(self, i):
        return (self.sample, self.img_dif_conditions)[i]


class UNet2DConditionModel(nn.Module):
    """Drop-in for the reference's `model.unet_2d_condition.UNet2DConditionModel` (same constructor
    :84-117, same forward signature :338-347, same 798 state-dict keys), executed by hand-written
    sm_100a kernels (engine.UNetEngine).  In train() mode with a trainable parameter on the path the forward is
    autograd-connected (storygen_b200/train.py: attn3 core forward / backward on the kernels, frozen layers as torch ops).
    """

    config_name = "config.json"

    def __init__(
        self,
        sample_size: Optional[int] = None,
        in_channels: int = 4,
        out_channels: int = 4,
        center_input_sample: bool = False,
        flip_sin_to_cos: bool = True,
        freq_shift: int = 0,
        down_block_types: Tuple[str, ...] = ("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D",
                                             "DownBlock2D"),
        mid_block_type: str = "UNetMidBlock2DCrossAttn",
        up_block_types: Tuple[str, ...] = ("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
        only_cross_attention: Union[bool, Tuple[bool]] = False,
        block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280),
        layers_per_block: int = 2,
        downsample_padding: int = 1,
        mid_block_scale_factor: float = 1,
        act_fn: str = "silu",
        norm_num_groups: Optional[int] = 32,
        norm_eps: float = 1e-5,
        cross_attention_dim: int = 1280,
        attention_head_dim: Union[int, Tuple[int]] = 8,
        dual_cross_attention: bool = False,
        use_linear_projection: bool = False,
        class_embed_type: Optional[str] = None,
        num_class_embeds: Optional[int] = None,
        upcast_attention: bool = False,
        resnet_time_scale_shift: str = "default",
        time_embedding_type: str = "positional",
        conv_in_kernel: int = 3,
        conv_out_kernel: int = 3,
        **unused,
    ):
        super().__init__()
        cfg = dict(sample_size=sample_size, in_channels=in_channels, out_channels=out_channels,
                   center_input_sample=center_input_sample, flip_sin_to_cos=flip_sin_to_cos, freq_shift=freq_shift,
                   down_block_types=tuple(down_block_types), mid_block_type=mid_block_type,
                   up_block_types=tuple(up_block_types), only_cross_attention=only_cross_attention,
                   block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
                   downsample_padding=downsample_padding, mid_block_scale_factor=mid_block_scale_factor, act_fn=act_fn,
                   norm_num_groups=norm_num_groups, norm_eps=norm_eps, cross_attention_dim=cross_attention_dim,
                   attention_head_dim=attention_head_dim, dual_cross_attention=dual_cross_attention,
                   use_linear_projection=use_linear_projection, class_embed_type=class_embed_type,
                   num_class_embeds=num_class_embeds, upcast_attention=upcast_attention,
                   resnet_time_scale_shift=resnet_time_scale_shift, time_embedding_type=time_embedding_type,
                   conv_in_kernel=conv_in_kernel, conv_out_kernel=conv_out_kernel)
        # the configurations the hand-written path implements (everything SD-1.5 / StoryGen uses)
        unsupported = []
        if act_fn not in ("silu", "swish"): unsupported.append("act_fn")
        if not flip_sin_to_cos or freq_shift != 0: unsupported.append("flip_sin_to_cos/freq_shift")
        if use_linear_projection or dual_cross_attention or upcast_attention: unsupported.append("projection/attention variant")
        if class_embed_type is not None or num_class_embeds is not None: unsupported.append("class embedding")
        if time_embedding_type != "positional" or resnet_time_scale_shift != "default": unsupported.append("time embedding variant")
        if conv_in_kernel != 3 or conv_out_kernel != 3 or downsample_padding != 1: unsupported.append("conv kernel/padding")
        if only_cross_attention not in (False, (False,) * len(down_block_types)): unsupported.append("only_cross_attention")
        if mid_block_type != "UNetMidBlock2DCrossAttn" or mid_block_scale_factor != 1: unsupported.append("mid block")
        if not isinstance(attention_head_dim, int): unsupported.append("per-block attention_head_dim")
        if unsupported:
            raise NotImplementedError(f"UNet2DConditionModel (B200 path): unsupported config: {unsupported}")
        self.config = _Config(cfg)
        for k, v in cfg.items():  # diffusers 0.13 register_to_config also sets attributes (unet.in_channels)
            setattr(self, k, v)

        boc, heads, xdim, groups, eps = tuple(block_out_channels), attention_head_dim, cross_attention_dim, norm_num_groups, norm_eps
        temb = boc[0] * 4
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], temb)
        self.down_blocks = nn.ModuleList()
        self.up_blocks = nn.ModuleList()
        out = boc[0]
        for i, t in enumerate(down_block_types):
            cin, out = out, boc[i]
            last = i == len(boc) - 1
            if t == "CrossAttnDownBlock2D":
                self.down_blocks.append(CrossAttnDownBlock2D(cin, out, temb, layers_per_block, eps, groups, heads, xdim, not last))
            elif t == "DownBlock2D":
                self.down_blocks.append(DownBlock2D(cin, out, temb, layers_per_block, eps, groups, not last))
            else:
                raise ValueError(f"{t} does not exist.")
        self.mid_block = UNetMidBlock2DCrossAttn(boc[-1], temb, eps, groups, heads, xdim)
        rev = list(reversed(boc))
        out = rev[0]
        self.num_upsamplers = 0
        for i, t in enumerate(up_block_types):
            prev, out = out, rev[i]
            cin = rev[min(i + 1, len(boc) - 1)]
            last = i == len(boc) - 1
            if not last:
                self.num_upsamplers += 1
            if t == "CrossAttnUpBlock2D":
                self.up_blocks.append(CrossAttnUpBlock2D(cin, out, prev, temb, layers_per_block + 1, eps, groups, heads, xdim, not last))
            elif t == "UpBlock2D":
                self.up_blocks.append(UpBlock2D(cin, out, prev, temb, layers_per_block + 1, eps, groups, not last))
            else:
                raise ValueError(f"{t} does not exist.")
        self.conv_norm_out = nn.GroupNorm(groups, boc[0], eps=eps)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)
        self._engine = None
        self.register_load_state_dict_post_hook(lambda m, k: m._invalidate())

    # ---- nn.Module plumbing -------------------------------------------------------------------
    def _invalidate(self):
        self._engine = None

    def _apply(self, fn, *a, **k):
        self._engine = None  # .to() / .cuda() / .half() move or re-type the weights
        return super()._apply(fn, *a, **k)

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device

    def refresh_weights(self):
        """Re-pack the kernel-side weight copies after parameters were modified in place."""
        self._engine = None

    # accepted for drop-in compatibility with the reference's callers (inference.py:58-64 etc.)
    def set_use_memory_efficient_attention_xformers(self, *a, **k):
        return None

    def enable_xformers_memory_efficient_attention(self, *a, **k):
        return None

    def set_attention_slice(self, *a, **k):
        return None

    def _set_gradient_checkpointing(self, *a, **k):
        return None

    # ---- the forward --------------------------------------------------------------------------
    def engine(self):
        if self._engine is None:
            from .engine import UNetEngine
            self._engine = UNetEngine(self)
        return self._engine

    def train(self, mode: bool = True):
        if mode != self.training:
            self._engine = None  # weights may have been updated by an optimiser: re-pack before the next inference call
        return super().train(mode)

    def _needs_autograd(self, image_hidden_states) -> bool:
        """train() mode with gradients enabled and a trainable parameter on the path: stage 2 trains attn3 only
        (train_StorySalon_stage2.py:172-177), which the feature passes (no image_hidden_states) never touch."""
        if not (self.training and torch.is_grad_enabled()):
            return False
        if image_hidden_states is not None:
            return any(p.requires_grad for p in self.parameters())
        return any(p.requires_grad for n, p in self.named_parameters() if ".attn3." not in n)

    def forward(self, sample: torch.Tensor, timestep: Union[torch.Tensor, float, int],
                encoder_hidden_states: torch.Tensor, image_hidden_states: Optional[Dict[str, torch.Tensor]] = None,
                class_labels: Optional[torch.Tensor] = None, cross_attention_kwargs: Optional[Dict[str, Any]] = None,
                return_dict: bool = True):
        """model/unet_2d_condition.py:338-485.  `sample` (N,C,H,W); `timestep` scalar, 0-d or (N,);
        `encoder_hidden_states` (N,L,D); `image_hidden_states` None (feature pass: returns the 16 taps
        (N, H_k W_k, C_k)) or a mapping with those keys holding (N, R H_k W_k, C_k) (denoise pass:
        returns an empty dict)."""
        if class_labels is not None:
            raise NotImplementedError("class conditioning is not part of the StoryGen configuration")
        if self.device.type != "cuda":
            raise RuntimeError("storygen_b200.UNet2DConditionModel runs only on a CUDA (sm_100a) device: move the "
                               "model with .to('cuda') first (there is no CPU fallback)")
        if self._needs_autograd(image_hidden_states):
            from .train import unet_forward_train
            out, taps = unet_forward_train(self, sample, timestep, encoder_hidden_states, image_hidden_states)
            if not return_dict:
                return (out, taps)
            return UNet2DConditionOutput(sample=out, img_dif_conditions=taps)
        with torch.no_grad():
            out, taps = self.engine().forward(sample, timestep, encoder_hidden_states, image_hidden_states)
        if self.dtype in (torch.float16, torch.bfloat16):
            out = out.to(self.dtype)  # the reference returns the sample in the weights' dtype (inference.py:72-74)
        if not return_dict:
            return (out, taps)
        return UNet2DConditionOutput(sample=out, img_dif_conditions=taps)

    # ---- checkpoint helpers -------------------------------------------------------------------
    def load_SDM_state_dict(self, state_dict_SDM, **kwargs):
        """model/unet_2d_condition.py:487-510: load an SD-1.5 UNet; attn3 <- attn1, norm4 <- norm1."""
        own = self.state_dict()
        sd = {}
        for k, v in state_dict_SDM.items():
            if k not in own:
                raise KeyError(f"SDM state_dict key {k} does not exist in model")
            if v.shape == own[k].shape:
                sd[k] = v
        for k in own:
            if k not in sd:
                sd[k] = sd[k.replace("attn3", "attn1").replace("norm4", "norm1")]
        self.load_state_dict(sd, **kwargs)

    @classmethod
    def from_config(cls, path_or_dict, subfolder: Optional[str] = None, **kwargs):
        if isinstance(path_or_dict, (str, os.PathLike)):
            p = os.path.join(path_or_dict, subfolder) if subfolder else path_or_dict
            if os.path.isdir(p):
                p = os.path.join(p, cls.config_name)
            with open(p) as f:
                cfg = json.load(f)
        else:
            cfg = dict(path_or_dict)
        cfg = {k: v for k, v in cfg.items() if not k.startswith("_")}
        cfg.update(kwargs)
        return cls(**cfg)

    @classmethod
    def from_pretrained(cls, path, subfolder: Optional[str] = None, torch_dtype=None, **kwargs):
        """diffusers directory layout: <path>/<subfolder>/config.json + diffusion_pytorch_model.{safetensors,bin}."""
        d = os.path.join(path, subfolder) if subfolder else path
        model = cls.from_config(d)
        st = os.path.join(d, "diffusion_pytorch_model.safetensors")
        if os.path.exists(st):
            from safetensors.torch import load_file
            sd = load_file(st)
        else:
            sd = torch.load(os.path.join(d, "diffusion_pytorch_model.bin"), map_location="cpu")
        model.load_state_dict(sd)
        if torch_dtype is not None:
            model = model.to(torch_dtype)
        return model.eval()

    def save_pretrained(self, save_directory, **kwargs):
        os.makedirs(save_directory, exist_ok=True)
        cfg = {k: (list(v) if isinstance(v, tuple) else v) for k, v in self.config.items()}
        cfg["_class_name"] = "UNet2DConditionModel"
        with open(os.path.join(save_directory, self.config_name), "w") as f:
            json.dump(cfg, f, indent=2)
        torch.save(self.state_dict(), os.path.join(save_directory, "diffusion_pytorch_model.bin"))
