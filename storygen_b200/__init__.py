"""storygen_b200 — B200-native (sm_100a) implementation of StoryGen's denoising hot path.

Public surface mirrors the reference (haoningwu3639/StoryGen):
    UNet2DConditionModel      <- model/unet_2d_condition.py
    StableDiffusionPipeline   <- model/pipeline.py
    DDIMScheduler             <- diffusers.schedulers.DDIMScheduler as configured by the reference
The arithmetic runs in hand-written CUDA (storygen_b200/csrc, C ABI in include/storygen_b200.h).
"""
from .modeling import UNet2DConditionModel, UNet2DConditionOutput  # noqa: F401
from .scheduler import DDIMScheduler  # noqa: F401
from .pipeline import StableDiffusionPipeline, StableDiffusionPipelineOutput  # noqa: F401

__all__ = ["UNet2DConditionModel", "UNet2DConditionOutput", "DDIMScheduler", "StableDiffusionPipeline",
           "StableDiffusionPipelineOutput"]
