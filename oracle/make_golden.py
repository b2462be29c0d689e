"""ORACLE (test infrastructure): generate tests/golden/* from the REFERENCE's own code.

Imports /root/reference/model/*.py unmodified (through oracle/ref_shim, the test-only stand-in for the
absent diffusers==0.13.1), loads the key-seeded synthetic weights of oracle/weights.py, runs

  * UNet2DConditionModel.forward — feature pass and R=2 denoise pass, tiny SD-shaped config,
  * StableDiffusionPipeline.__call__ — 2 DDIM steps, 2 prior frames, stub tokenizer/CLIP/VAE,
  * DDIMScheduler tables,
  * the 798-entry state-dict schema of the full SD-1.5 + attn3 configuration,

and stores inputs' seeds + outputs as small fixtures.  /root/reference does not exist on the GPU box;
the fixtures (and this script) are what travels.   Usage:  python -m oracle.make_golden
"""
from __future__ import annotations

import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"


def import_reference():
    if not os.path.isdir(REF):
        raise RuntimeError(f"{REF} is not mounted")
    for p in (os.path.join(ROOT, "oracle", "ref_shim"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    from model.unet_2d_condition import UNet2DConditionModel  # noqa
    from model.pipeline import StableDiffusionPipeline  # noqa
    from diffusers.schedulers import DDIMScheduler  # noqa  (shim)
    return UNet2DConditionModel, StableDiffusionPipeline, DDIMScheduler


def sub(t: torch.Tensor, n: int = 4096) -> torch.Tensor:
    """strided subsample (keeps fixtures small)"""
    f = t.reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step][:n].clone()


def unet_inputs(n: int, xdim: int, hw: int = 64, seed: int = 11):
    g = torch.Generator("cpu").manual_seed(seed)
    x = torch.randn(n, 4, hw, hw, generator=g)
    txt = torch.randn(n, 77, xdim, generator=g)
    t = torch.tensor([98, 294, 981, 5, 640, 33][:n])
    return x, t, txt


def cond_from_taps(taps, R):
    return {k: torch.cat([v * (1.0 + 0.1 * r) for r in range(R)], dim=1) for k, v in taps.items()}


PIPE_ARGS = dict(stage="auto-regressive", prompt=["a dog runs on the grass"], prev_prompt=[["a dog sits"], ["the dog sleeps now"]],
                 height=512, width=512, num_inference_steps=2, guidance_scale=7.0, image_guidance_scale=3.5,
                 num_images_per_prompt=1)


def pipeline_images(seed=5):
    g = torch.Generator("cpu").manual_seed(seed)
    return torch.rand(1, 2, 3, 512, 512, generator=g)


def main():
    from oracle.stubs import StubTextEncoder, StubTokenizer, StubVAE
    from oracle.unet import SD15_CONFIG, TINY_CONFIG
    from oracle.weights import load_seeded

    UNet, Pipe, DDIM = import_reference()
    os.makedirs(GOLD, exist_ok=True)
    torch.set_grad_enabled(False)

    # ---- state-dict schema of the full configuration ------------------------------------------
    with torch.device("meta"):
        full = UNet(**{k: v for k, v in SD15_CONFIG.items()})
    schema = {k: list(v.shape) for k, v in full.state_dict().items()}
    with open(os.path.join(GOLD, "state_dict_schema.json"), "w") as f:
        json.dump(schema, f)
    print("schema:", len(schema), "tensors,", sum(torch.Size(s).numel() for s in schema.values()), "params")

    # ---- UNet forward, tiny config ---------------------------------------------------------------
    ref = load_seeded(UNet(**TINY_CONFIG).eval(), seed=0)
    x, t, txt = unet_inputs(2, TINY_CONFIG["cross_attention_dim"])
    s_feat, taps = ref(x, t, txt, return_dict=False)
    cond = cond_from_taps(taps, 2)
    s_den, d = ref(x, 981, txt, image_hidden_states=cond, return_dict=False)
    assert len(d) == 0 and len(taps) == 16
    torch.save(dict(config=TINY_CONFIG, weight_seed=0, input_seed=11, n=2, R=2,
                    feature_sample=s_feat.clone(), denoise_sample=s_den.clone(),
                    taps_sub={k: sub(v) for k, v in taps.items()},
                    taps_shape={k: list(v.shape) for k, v in taps.items()}),
               os.path.join(GOLD, "unet_tiny.pt"))
    print("unet_tiny: feature std %.4f denoise std %.4f" % (s_feat.std(), s_den.std()))

    # ---- scheduler tables ------------------------------------------------------------------------
    sch = DDIM(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", num_train_timesteps=1000,
               set_alpha_to_one=False, steps_offset=1, clip_sample=False)
    out = {}
    for T in (50, 40, 4):
        sch.set_timesteps(T)
        out[f"timesteps_{T}"] = sch.timesteps.tolist()
    out["alphas_cumprod_idx"] = [0, 1, 98, 294, 500, 981, 999]
    out["alphas_cumprod"] = [float(sch.alphas_cumprod[i]) for i in out["alphas_cumprod_idx"]]
    g = torch.Generator("cpu").manual_seed(3)
    xs, es = torch.randn(1, 4, 8, 8, generator=g), torch.randn(1, 4, 8, 8, generator=g)
    sch.set_timesteps(50)
    out["step_981"] = sch.step(es, 981, xs, eta=0.0).prev_sample.flatten().tolist()
    out["step_1"] = sch.step(es, 1, xs, eta=0.0).prev_sample.flatten().tolist()
    out["add_noise_294"] = sch.add_noise(xs, es, torch.tensor([294])).flatten().tolist()
    with open(os.path.join(GOLD, "scheduler.json"), "w") as f:
        json.dump(out, f)

    # ---- full pipeline call with stubs, both stages that condition on prior frames -------------------
    # (model/pipeline.py:419-436: the stages differ only in the timesteps the frames are noised to)
    for stage, fname in (("auto-regressive", "pipeline_tiny.pt"), ("multi-image-condition", "pipeline_tiny_mic.pt")):
        sch2 = DDIM(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", num_train_timesteps=1000,
                    set_alpha_to_one=False, steps_offset=1, clip_sample=False)
        pipe = Pipe(vae=StubVAE(), text_encoder=StubTextEncoder(TINY_CONFIG["cross_attention_dim"]),
                    tokenizer=StubTokenizer(), unet=ref, scheduler=sch2)
        lat_steps = []
        torch.manual_seed(1234)  # global RNG: VAE posterior samples + the shared `noise` tensor
        gen = torch.Generator("cpu").manual_seed(99)
        pipe(image_prompt=pipeline_images(), generator=gen, output_type="np",
             callback=lambda i, t, lat: lat_steps.append(lat.clone()), **dict(PIPE_ARGS, stage=stage))
        assert len(lat_steps) == 2, len(lat_steps)
        torch.save(dict(latents_step=[l.clone() for l in lat_steps], global_seed=1234, gen_seed=99, image_seed=5,
                        text_seed=7, stage=stage), os.path.join(GOLD, fname))
        print("%s (%s): final latent std %.4f" % (fname, stage, lat_steps[-1].std()))
    for fn in sorted(os.listdir(GOLD)):
        print("  ", fn, os.path.getsize(os.path.join(GOLD, fn)), "bytes")


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    main()
