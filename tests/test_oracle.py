"""CPU: the oracle restatement against (a) the committed golden fixtures generated from the reference's
own code (oracle/make_golden.py) and (b) — where /root/reference is mounted — the reference itself,
imported unmodified through oracle/ref_shim."""
import json
import os

import pytest
import torch

from conftest import GOLD, have_reference
from oracle import sampler
from oracle.make_golden import PIPE_ARGS, cond_from_taps, pipeline_images, sub, unet_inputs
from oracle.stubs import StubTextEncoder, StubTokenizer, StubVAE
from oracle.unet import SD15_CONFIG, TINY_CONFIG, OracleUNet, timestep_features
from oracle.weights import load_seeded

torch.set_grad_enabled(False)


@pytest.fixture(scope="module")
def tiny_oracle():
    return load_seeded(OracleUNet(**TINY_CONFIG).eval(), seed=0)


def test_state_dict_schema_matches_reference():
    with open(os.path.join(GOLD, "state_dict_schema.json")) as f:
        schema = json.load(f)
    with torch.device("meta"):
        o = OracleUNet(**SD15_CONFIG)
    sd = o.state_dict()
    assert len(schema) == 798 and set(schema) == set(sd)
    assert all(list(sd[k].shape) == schema[k] for k in schema)
    assert sum(v.numel() for v in sd.values()) == 909120004


def test_unet_forward_matches_reference_golden(tiny_oracle):
    g = torch.load(os.path.join(GOLD, "unet_tiny.pt"))
    x, t, txt = unet_inputs(g["n"], TINY_CONFIG["cross_attention_dim"], seed=g["input_seed"])
    s, taps = tiny_oracle(x, t, txt)
    assert sorted(taps) == sorted(g["taps_sub"]) and len(taps) == 16
    assert (s - g["feature_sample"]).abs().max() < 2e-4
    for k, v in taps.items():
        assert list(v.shape) == g["taps_shape"][k]
        assert (sub(v) - g["taps_sub"][k]).abs().max() < 2e-4, k
    s2, d = tiny_oracle(x, 981, txt, cond_from_taps(taps, g["R"]))
    assert len(d) == 0
    assert (s2 - g["denoise_sample"]).abs().max() < 2e-4


def test_scheduler_matches_golden():
    with open(os.path.join(GOLD, "scheduler.json")) as f:
        g = json.load(f)
    for T in (50, 40, 4):
        assert sampler.ddim_timesteps(T) == g[f"timesteps_{T}"]
    acp = sampler.alphas_cumprod()
    for i, v in zip(g["alphas_cumprod_idx"], g["alphas_cumprod"]):
        assert abs(float(acp[i]) - v) < 1e-7
    gen = torch.Generator("cpu").manual_seed(3)
    xs, es = torch.randn(1, 4, 8, 8, generator=gen), torch.randn(1, 4, 8, 8, generator=gen)
    assert torch.allclose(sampler.ddim_step(es, 981, xs, 50, acp).flatten(), torch.tensor(g["step_981"]), atol=1e-5)
    assert torch.allclose(sampler.ddim_step(es, 1, xs, 50, acp).flatten(), torch.tensor(g["step_1"]), atol=1e-5)
    assert torch.allclose(sampler.add_noise(xs, es, torch.tensor([294]), acp).flatten(),
                          torch.tensor(g["add_noise_294"]), atol=1e-6)
    assert sampler.feature_timesteps(981, 3, "auto-regressive") == [294, 196, 98]
    assert sampler.feature_timesteps(21, 3, "auto-regressive") == [6, 4, 2]
    assert sampler.feature_timesteps(981, 3, "multi-image-condition") == [98, 98, 98]


def test_timestep_features_order():
    f = timestep_features(torch.tensor([0, 1, 981]), 320)
    assert f.shape == (3, 320)
    assert torch.allclose(f[0, :160], torch.ones(160)) and torch.allclose(f[0, 160:], torch.zeros(160))  # [cos | sin]
    assert abs(float(f[1, 160]) - 0.8414709848) < 1e-6  # sin(1 * 10000^0)


def stub_pipeline_inputs(g, xdim):
    """Re-derive, with the same stubs / seeds / RNG order as the reference pipeline run in make_golden.py,
    the tensors that reach the denoising loop."""
    tok, enc, vae = StubTokenizer(), StubTextEncoder(xdim, seed=g["text_seed"]), StubVAE()
    emb = lambda p: enc(tok(p, max_length=77).input_ids)[0]
    text_cond, text_uncond = emb(PIPE_ARGS["prompt"]), emb([""])
    prev_cond = [emb(p) for p in PIPE_ARGS["prev_prompt"]]
    prev_uncond = [text_uncond for _ in prev_cond]
    gen = torch.Generator("cpu").manual_seed(g["gen_seed"])
    latents = torch.randn(1, 4, 64, 64, generator=gen)
    torch.manual_seed(g["global_seed"])
    frames = torch.transpose(pipeline_images(g["image_seed"]), 0, 1)
    zero = vae.encode(frames[0] * 0).latent_dist.sample() * 0.18215
    frame_latents = [vae.encode(fr).latent_dist.sample() * 0.18215 for fr in frames]
    noise = torch.randn_like(frame_latents[0])
    return dict(latents=latents, text_uncond=text_uncond, text_cond=text_cond, prev_uncond=prev_uncond,
                prev_cond=prev_cond, frame_latents=frame_latents, zero_latents=zero, noise=noise)


@pytest.mark.parametrize("fname,stage", [("pipeline_tiny.pt", "auto-regressive"),
                                         ("pipeline_tiny_mic.pt", "multi-image-condition")])
def test_denoise_loop_matches_reference_pipeline_golden(tiny_oracle, fname, stage):
    """Both frame-conditioned stages of the reference's StableDiffusionPipeline.__call__ (model/pipeline.py:419-436)."""
    g = torch.load(os.path.join(GOLD, fname))
    assert g["stage"] == stage
    inp = stub_pipeline_inputs(g, TINY_CONFIG["cross_attention_dim"])
    steps = []
    out = sampler.denoise(tiny_oracle, num_inference_steps=PIPE_ARGS["num_inference_steps"],
                          guidance_scale=PIPE_ARGS["guidance_scale"], image_guidance_scale=PIPE_ARGS["image_guidance_scale"],
                          stage=stage, callback=lambda t, lat: steps.append(lat.clone()), **inp)
    assert len(steps) == len(g["latents_step"])
    for a, b in zip(steps, g["latents_step"]):
        assert (a - b).abs().max() < 5e-4 * max(1.0, float(b.abs().max()))
    assert torch.equal(out, steps[-1])


@pytest.mark.skipif(not have_reference(), reason="/root/reference is not mounted (GPU box)")
def test_oracle_equals_reference_live():
    """Reference model/*.py imported unmodified (ref_shim) vs the oracle, fresh random inputs."""
    from oracle.make_golden import import_reference
    UNet, _, _ = import_reference()
    cfg = dict(TINY_CONFIG)
    ref = load_seeded(UNet(**cfg).eval(), seed=3)
    ora = OracleUNet(**cfg).eval()
    ora.load_state_dict(ref.state_dict())
    x, t, txt = unet_inputs(3, cfg["cross_attention_dim"], seed=21)
    rs, rt = ref(x, t, txt, return_dict=False)
    os_, ot = ora(x, t, txt)
    assert (rs - os_).abs().max() < 1e-4
    assert max((rt[k] - ot[k]).abs().max() for k in rt) < 1e-4
    cond = cond_from_taps(rt, 3)
    r2 = ref(x, t, txt, image_hidden_states=cond, return_dict=False)[0]
    o2 = ora(x, t, txt, cond)[0]
    assert (r2 - o2).abs().max() < 1e-4
    # load_SDM_state_dict: attn3 <- attn1, norm4 <- norm1
    sdm = {k: v for k, v in ref.state_dict().items() if "attn3" not in k and "norm4" not in k}
    ora.load_SDM_state_dict(dict(sdm))
    ref.load_SDM_state_dict(dict(sdm))
    assert all(torch.equal(ora.state_dict()[k], ref.state_dict()[k]) for k in ref.state_dict())
