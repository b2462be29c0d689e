"""GPU: parity on the configurations the BENCHMARK actually runs ("What's weak" #1) — full-size SD-1.5
UNet + attn3, the 50-step dual-guidance trajectory at guidance 7.0 / 3.5 through `pipe.denoise` (batched feature
lanes, kv_map de-duplication, text-KV cache, CUDA-graph replay), the 768 px / 5-frame attention shapes, and the
fp16 hand-over of inference.py:35,131.

The fp32 reference here is the ORACLE run on the GPU in fp32 torch (TF32 off): same module as the CPU oracle that is
pinned to the reference's model/*.py (tests/test_oracle.py), just on a device where 550 TFLOP take seconds.

Gate of the trajectory test: the deviation of this path from the fp32 oracle must not exceed the
deviation of the ORACLE ITSELF under torch's bf16 autocast (what the reference's `mixed_precision="bf16"` run
does) on the same seeds, times a slack factor that covers the sampling noise of a max over 16 384 latent values of
two independent rounding histories (both numbers are printed; profiles/r02_parity_bench_config.log keeps them)."""
import os

import pytest
import torch

from oracle import sampler
from oracle.make_golden import cond_from_taps, unet_inputs
from oracle.unet import SD15_CONFIG, OracleUNet
from oracle.weights import load_seeded
from storygen_b200 import DDIMScheduler, StableDiffusionPipeline, UNet2DConditionModel

pytestmark = pytest.mark.gpu
DEV = "cuda"
SLACK_MAX, SLACK_RMS = 1.5, 1.25


def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


@pytest.fixture(scope="module")
def full():
    _no_tf32()
    prev = torch.is_grad_enabled()
    torch.set_grad_enabled(False)
    ora = load_seeded(OracleUNet(**SD15_CONFIG).eval(), 0).to(DEV)
    mod = UNet2DConditionModel(**SD15_CONFIG)
    mod.load_state_dict(ora.state_dict())
    mod = mod.to(DEV).eval()
    yield ora, mod
    del ora, mod
    torch.cuda.empty_cache()
    torch.set_grad_enabled(prev)


def _dev(inp):
    return {k: ([t.to(DEV) for t in v] if isinstance(v, list) else v.to(DEV)) for k, v in inp.items()}


def test_50_step_trajectory_full_size_vs_fp32_oracle_and_bf16_autocast(full):
    """cfg2's generation at B=1: 50 DDIM steps, 3 prior frames, guidance 7.0 / 3.5, CUDA-graph replay."""
    ora, mod = full
    inp = _dev(sampler.synthetic_inputs(batch=1, num_frames=3))
    kw = dict(num_inference_steps=50, guidance_scale=7.0, image_guidance_scale=3.5, stage="auto-regressive")
    every = {}

    def collect(name):
        every[name] = []
        return lambda *a: every[name].append(a[-1].float().clone())

    ref32 = sampler.denoise(ora, callback=collect("fp32"), **kw, **inp)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ref16 = sampler.denoise(ora, callback=collect("bf16"), **kw, **inp).float()
    pipe = StableDiffusionPipeline(None, None, None, mod, DDIMScheduler())
    ours = pipe.denoise(callback=collect("ours"), **kw, **inp)
    assert pipe.kernel_launches > 0 and torch.isfinite(ours).all()

    def dev(a, b):
        d = (a - b).abs()
        return d.max().item(), d.pow(2).mean().sqrt().item()

    std = ref32.std().item()
    o_max, o_rms = dev(ours, ref32)
    b_max, b_rms = dev(ref16, ref32)
    print(f"\n[trajectory 50 steps, full size, B=1 R=3, guidance 7/3.5] std(latent) {std:.4f}\n"
          f"  this path      vs fp32 oracle: max|d| {o_max:.4f}  rms {o_rms:.4f}\n"
          f"  oracle bf16 autocast vs fp32 : max|d| {b_max:.4f}  rms {b_rms:.4f}")
    for s in (0, 1, 9, 24, 49):
        om, _ = dev(every["ours"][s], every["fp32"][s])
        bm, _ = dev(every["bf16"][s], every["fp32"][s])
        print(f"  after step {s + 1:2d}: ours {om:.4f}   bf16-autocast {bm:.4f}")
    assert o_max <= SLACK_MAX * b_max, f"max|d latent| {o_max:.4g} > {SLACK_MAX} x bf16-autocast deviation {b_max:.4g}"
    assert o_rms <= SLACK_RMS * b_rms, f"rms d latent {o_rms:.4g} > {SLACK_RMS} x bf16-autocast deviation {b_rms:.4g}"


@pytest.mark.parametrize("hw,R", [(64, 5), (96, 5)])
def test_full_size_unet_768px_and_5_frames_vs_oracle(full, hw, R):
    """cfg5's extreme attention shapes on the full-size UNet: Nk = R * HW = 20 480 (512 px) and 46 080 (768 px; the
    reference itself mis-keys the cond dict off 512 px, SURVEY D4, so the oracle's block-index keying is the definition)."""
    ora, mod = full
    x, t, txt = unet_inputs(1, 768, hw=hw, seed=100 + hw)
    x, t, txt = x.to(DEV), t.to(DEV), txt.to(DEV)
    os_, ot = ora(x, t, txt)
    cs, ct = mod(x, t, txt, return_dict=False)
    rel = 0.10

    def check(got, ref, what):
        err, tol = (got.float() - ref.float()).abs().max().item(), rel * ref.float().std().item()
        assert err <= tol, f"{what}: max|d| {err:.4g} > {tol:.4g}"
        return err / ref.float().std().item()

    e = [check(cs, os_, "feature sample")] + [check(ct[k], ot[k], k) for k in ot]
    cond = cond_from_taps(ot, R)
    o2 = ora(x, 981, txt, cond)[0]
    c2 = mod(x, 981, txt, image_hidden_states=cond, return_dict=False)[0]
    e2 = check(c2, o2, f"denoise sample R={R}")
    print(f"\n[full size, latent {hw}, R={R}, Nk={R * hw * hw}] max|d|/std: feature {max(e):.3f}  denoise {e2:.3f}")


def test_fp16_handover_full_size(full):
    """inference.py:35,131: `unet.to(device, dtype=torch.float16)` and fp16 latents / embeddings.  This path computes in
    bf16 with fp32 accumulation whatever dtype the caller hands over (fp16 weights are re-packed to bf16), so the stated
    tolerance is the bf16 one: max|d| <= 0.10 x std against the fp32 oracle holding the SAME fp16-rounded weights
    (the reference's own fp16 autocast deviates by 0.7 % of std, its bf16 autocast by 5.9 %: SURVEY.md §8c).
    The returned sample has the model's dtype, as the reference's does."""
    ora, mod = full
    sd16 = {k: v.half() for k, v in ora.state_dict().items()}
    ora16 = OracleUNet(**SD15_CONFIG).eval()
    ora16.load_state_dict({k: v.float() for k, v in sd16.items()})
    ora16 = ora16.to(DEV)
    m16 = UNet2DConditionModel(**SD15_CONFIG)
    m16.load_state_dict(sd16)
    m16 = m16.to(DEV, dtype=torch.float16).eval()
    assert m16.dtype == torch.float16
    x, t, txt = unet_inputs(1, 768, seed=78)
    x, txt = x.half(), txt.half()
    os_, ot = ora16(x.float().to(DEV), t.to(DEV), txt.float().to(DEV))
    cs, ct = m16(x.to(DEV), t.to(DEV), txt.to(DEV), return_dict=False)
    assert cs.dtype == torch.float16
    err, std = (cs.float() - os_).abs().max().item(), os_.std().item()
    assert err <= 0.10 * std, f"fp16 hand-over: max|d| {err:.4g} > 0.10 x std {std:.4g}"
    cond = cond_from_taps(ot, 3)
    o2 = ora16(x.float().to(DEV), 981, txt.float().to(DEV), cond)[0]
    c2 = m16(x.to(DEV), 981, txt.to(DEV), image_hidden_states={k: v.half() for k, v in cond.items()}, return_dict=False)[0]
    err2 = (c2.float() - o2).abs().max().item()
    assert err2 <= 0.10 * o2.std().item()
    print(f"\n[fp16 hand-over, full size] max|d|/std: feature {err / std:.3f}  denoise {err2 / o2.std().item():.3f}")
    del ora16, m16
    torch.cuda.empty_cache()


def test_bench_weights_and_fixture_match_the_oracle_recipe():
    """bench.py checks its own output against tests/golden/bench_steps_fp32.pt; the fixture is only meaningful if the
    product-side name-seeded init equals the oracle's (same tensors for the same key and seed)."""
    from oracle.unet import TINY_CONFIG
    from storygen_b200.synthetic import seeded_init_
    a = load_seeded(OracleUNet(**TINY_CONFIG), 0).state_dict()
    b = seeded_init_(UNet2DConditionModel(**TINY_CONFIG), 0).state_dict()
    assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)
    p = os.path.join(os.path.dirname(__file__), "golden", "bench_steps_fp32.pt")
    if os.path.exists(p):
        g = torch.load(p)
        assert g["weights_seed"] == 0 and len(g["latents_step"]) >= 2 and g["latents_step"][0].shape[1:] == (4, 64, 64)
