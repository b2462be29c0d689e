"""CPU: host-side logic of the product package and the C-ABI surface (no kernel is launched)."""
import ctypes
import json
import os
import re

import pytest
import torch

from conftest import GOLD, ROOT
from storygen_b200 import DDIMScheduler, StableDiffusionPipeline, UNet2DConditionModel, _lib, ops
from storygen_b200.pipeline import feature_timesteps

TINY = dict(block_out_channels=(64, 128, 256, 256), norm_num_groups=8, cross_attention_dim=64, attention_head_dim=2)


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "storygen_b200.h")).read()
    declared = set(re.findall(r"\b(sg_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.sg_abi_version() == _lib.SG_ABI_VERSION
    assert lib.sg_sizeof_gemm_args() == ctypes.sizeof(_lib.SgGemmArgs)
    assert lib.sg_sizeof_attn_args() == ctypes.sizeof(_lib.SgAttnArgs)
    assert lib.sg_error_string(-1).decode().startswith("unsupported")


def test_argument_validation_without_gpu():
    lib = _lib.load()
    assert lib.sg_gemm(None, None) == -4  # SG_ERR_ARG
    a = _lib.SgGemmArgs()
    a.b, a.out, a.num_src, a.num_segs = 16, 16, 1, 1
    a.bw, a.bh, a.bn = 64, 1, 1  # not 128 rows
    assert lib.sg_gemm(ctypes.byref(a), None) == -1  # SG_ERR_SHAPE
    at = _lib.SgAttnArgs()
    at.q = at.k = at.v = at.o = 16
    at.nq = at.nk = at.n_lanes = at.n_heads = 1
    at.head_dim = 48
    assert lib.sg_attention(ctypes.byref(at), None) == -1


def test_tile_box():
    for w, h, n in [(64, 64, 12), (32, 32, 3), (16, 16, 1), (8, 8, 5), (96, 96, 2), (48, 48, 2), (24, 24, 3), (12, 12, 8),
                    (4096, 1, 1), (300, 1, 1)]:
        bw, bh, bn = ops.tile_box(w, h, n)
        assert bw * bh * bn == 128 and bw <= 256 and bh <= 256 and bn <= 256


def test_state_dict_schema_and_constructor_surface():
    with open(os.path.join(GOLD, "state_dict_schema.json")) as f:
        schema = json.load(f)
    with torch.device("meta"):
        m = UNet2DConditionModel(cross_attention_dim=768)
    sd = m.state_dict()
    assert set(sd) == set(schema) and all(list(sd[k].shape) == schema[k] for k in schema)
    # what the reference's callers touch
    assert m.in_channels == 4 and m.config.sample_size is None and m.config["cross_attention_dim"] == 768
    trainable = [n for n, _ in m.named_modules() if n.endswith("attn3")]
    assert len(trainable) == 16
    assert sum(p.numel() for n, mod in m.named_modules() if n.endswith("attn3") for p in mod.parameters()) == 49574080
    m.set_use_memory_efficient_attention_xformers(True)
    m.set_attention_slice("auto")
    with pytest.raises(NotImplementedError):
        UNet2DConditionModel(use_linear_projection=True)
    with pytest.raises(NotImplementedError):
        UNet2DConditionModel(class_embed_type="timestep")


def test_no_cpu_fallback():
    m = UNet2DConditionModel(**TINY)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 4, 64, 64), 1, torch.zeros(1, 77, 64))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.im2col3x3_nchw(torch.zeros(1, 4, 8, 8), torch.zeros(64, 64, dtype=torch.bfloat16))


def test_load_sdm_state_dict_and_save_roundtrip(tmp_path):
    m = UNet2DConditionModel(**TINY)
    sdm = {k: torch.randn_like(v) for k, v in m.state_dict().items() if "attn3" not in k and "norm4" not in k}
    m.load_SDM_state_dict(dict(sdm))
    sd = m.state_dict()
    for k in sd:
        src = k.replace("attn3", "attn1").replace("norm4", "norm1")
        assert torch.equal(sd[k], sdm[src])
    with pytest.raises(KeyError):
        m.load_SDM_state_dict({"not.a.key": torch.zeros(1)})
    m.save_pretrained(str(tmp_path / "unet"))
    m2 = UNet2DConditionModel.from_pretrained(str(tmp_path), subfolder="unet")
    assert all(torch.equal(m2.state_dict()[k], sd[k]) for k in sd)
    m3 = UNet2DConditionModel.from_config(str(tmp_path), subfolder="unet")
    assert m3.config.block_out_channels == (64, 128, 256, 256)


def test_scheduler_matches_golden():
    with open(os.path.join(GOLD, "scheduler.json")) as f:
        g = json.load(f)
    s = DDIMScheduler()
    for T in (50, 40, 4):
        s.set_timesteps(T)
        assert s.timesteps.tolist() == g[f"timesteps_{T}"]
    for i, v in zip(g["alphas_cumprod_idx"], g["alphas_cumprod"]):
        assert abs(float(s.alphas_cumprod[i]) - v) < 1e-7
    gen = torch.Generator("cpu").manual_seed(3)
    xs, es = torch.randn(1, 4, 8, 8, generator=gen), torch.randn(1, 4, 8, 8, generator=gen)
    s.set_timesteps(50)
    assert torch.allclose(s.step(es, 981, xs).prev_sample.flatten(), torch.tensor(g["step_981"]), atol=1e-5)
    assert torch.allclose(s.step(es, 1, xs).prev_sample.flatten(), torch.tensor(g["step_1"]), atol=1e-5)
    assert torch.allclose(s.add_noise(xs, es, torch.tensor([294])).flatten(), torch.tensor(g["add_noise_294"]), atol=1e-6)
    a_t, a_prev = s.step_alphas(1)
    assert a_prev == float(s.alphas_cumprod[0])  # set_alpha_to_one=False
    ref_cfg = "/root/reference/ckpt/stable-diffusion-v1-5"
    if os.path.isdir(ref_cfg):
        s2 = DDIMScheduler.from_pretrained(ref_cfg, subfolder="scheduler")
        assert torch.equal(s2.alphas_cumprod, s.alphas_cumprod)


def test_feature_timesteps_and_pipeline_surface():
    assert feature_timesteps(981, 3, "auto-regressive") == [294, 196, 98]
    assert feature_timesteps(1, 3, "auto-regressive") == [0, 0, 0]
    assert feature_timesteps(961, 5, "multi-image-condition") == [96] * 5
    with pytest.raises(NotImplementedError):
        feature_timesteps(1, 1, "no")
    import inspect
    sig = inspect.signature(StableDiffusionPipeline.__call__)
    assert list(sig.parameters)[1:] == ["stage", "prompt", "image_prompt", "prev_prompt", "height", "width",
                                        "num_inference_steps", "guidance_scale", "image_guidance_scale", "negative_prompt",
                                        "num_images_per_prompt", "eta", "generator", "latents", "output_type", "return_dict",
                                        "callback", "callback_steps"]
    fsig = inspect.signature(UNet2DConditionModel.forward)
    assert list(fsig.parameters)[1:] == ["sample", "timestep", "encoder_hidden_states", "image_hidden_states",
                                         "class_labels", "cross_attention_kwargs", "return_dict"]


def test_numpy_to_pil_returns_one_list_per_image():
    """model/pipeline.py:486-491: `.images` is a list of per-image lists; inference.py:117-120 saves `image[0]`."""
    import numpy as np

    from storygen_b200 import StableDiffusionPipeline
    imgs = StableDiffusionPipeline.numpy_to_pil(np.random.rand(3, 16, 16, 3).astype("float32"))
    assert isinstance(imgs, list) and len(imgs) == 3
    assert all(isinstance(i, list) and len(i) == 1 and i[0].size == (16, 16) for i in imgs)
    one = StableDiffusionPipeline.numpy_to_pil(np.random.rand(16, 16, 3).astype("float32"))
    assert len(one) == 1 and one[0][0].mode == "RGB"
