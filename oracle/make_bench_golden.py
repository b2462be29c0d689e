"""ORACLE (test infrastructure): fp32 reference of the first DDIM steps of the BENCH workload.

bench.py cannot run the fp32 oracle for a whole 50-step batch inside its timed budget, and /root/reference does
not exist on the GPU box, so the check it prints (`parity.max_abs_delta_step2`) compares the latents the CUDA path
holds after the first 2 of the 50 DDIM steps (guidance 7.0 / 3.5, 3 prior frames, 512 px, the seeded bench weights
and inputs) with this fixture:

    python -m oracle.make_bench_golden [--images 8] [--steps 2]   ->  tests/golden/bench_steps_fp32.pt

Images are keyed by their global index and are independent of each other, so one fixture of 8 images serves every
per-GPU batch size <= 8 of rank 0.  Runs the oracle (== the reference's model/*.py, tests/test_oracle.py) on the CPU.
"""
import argparse
import os
import time

import torch

from oracle import sampler
from oracle.unet import SD15_CONFIG, OracleUNet
from oracle.weights import load_seeded


class _Stop(Exception):
    pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=8)
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--latent", type=int, default=64)
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests",
                                                  "golden", "bench_steps_fp32.pt"))
    a = ap.parse_args()
    torch.set_grad_enabled(False)
    net = load_seeded(OracleUNet(**SD15_CONFIG).eval(), 0)
    inp = sampler.synthetic_inputs(a.images, a.frames, latent_hw=a.latent)
    lat = []

    def cb(t, latents):
        lat.append(latents.clone())
        print(f"step {len(lat)} (t={t}) done at {time.time() - t0:.0f} s", flush=True)
        if len(lat) == a.steps:
            raise _Stop

    t0 = time.time()
    try:
        sampler.denoise(net, num_inference_steps=50, guidance_scale=7.0, image_guidance_scale=3.5, stage="auto-regressive",
                        callback=cb, **inp)
    except _Stop:
        pass
    torch.save({"latents_step": [x.half().float() if False else x for x in lat], "images": a.images, "frames": a.frames,
                "latent": a.latent, "weights_seed": 0, "guidance": (7.0, 3.5), "ddim_steps": 50,
                "what": "fp32 oracle latents after DDIM step 1..n of the bench workload (oracle/make_bench_golden.py)"},
               a.out)
    print("wrote", a.out)


if __name__ == "__main__":
    main()
