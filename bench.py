#!/usr/bin/env python
"""Benchmark of StoryGen's denoising hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this framework (sm_100a kernels)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's algorithm on the host CPU cores

Workload (config.workload): SD-1.5-shaped UNet + StoryGen attn3 (909 M params, seeded random weights —
no checkpoint exists offline), 512x512 (latent 64x64), 3 prior frames, 50 DDIM steps, dual guidance,
bf16 storage / fp32 accumulation, B images per GPU ("cfg2", B=4).  ONE bench step = one full 50-step
generation of B images on every rank (weak scaling) followed by the single NCCL all-gather of the
final latents; value = images / second over all ranks.  Inputs are synthetic latents + embeddings
(tokenizer / CLIP / VAE are outside the measured path, as in the metric's definition).

Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
One step streams ~1.8 GB of weights and several GB of activations through a 126 MB L2 200 times, so
no explicit L2 flush is needed ("inputs larger than L2").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# algorithmic work of the reference path (2*MAC, per CFG lane, latent 64x64)
GF_FEATURE_PASS = 803.3
GF_DENOISE_PASS = {1: 976.9, 3: 1273.1, 5: 1569.3}


def tflop_per_image(R: int, steps: int) -> float:
    den = GF_DENOISE_PASS.get(R, 976.9 + (R - 1) * 148.1)
    return 3 * steps * (R * GF_FEATURE_PASS + den) / 1e3


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower() == "active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def usable_cores() -> int:
    """Cores this process may actually use (affinity mask and cgroup quota), not the box's core count."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period))))
    except Exception:
        pass
    return max(1, n)


class CpuReference:
    """The reference's algorithm on the host cores: the fp32 oracle (asserted equal to the reference's
    model/*.py in tests/test_oracle.py; the reference itself needs diffusers==0.13.1, absent offline).
    One SAMPLE = one feature-pass UNet call + one denoise-pass UNet call (R prior frames) for ONE CFG lane
    (0.80 + 1.27 TFLOP at R=3).  A denoising step of one image is 3 lanes x (R feature + 1 denoise) calls,
    every lane independent, so  t_step = 3 * (R * t_feature + t_denoise)  and images/s = 1 / (steps * t_step).
    Attention uses SDPA (same math as the default baddbmm path, without materialising 4.8 GB of scores)."""

    def __init__(self, R: int, threads: int):
        import torch

        from oracle import sampler
        from oracle.unet import SD15_CONFIG, OracleUNet
        self.torch, self.R = torch, R
        torch.set_num_threads(threads)
        torch.manual_seed(0)
        with torch.no_grad():
            self.net = OracleUNet(**SD15_CONFIG).eval()
        self.inp = sampler.synthetic_inputs(1, R)
        self.acp = sampler.alphas_cumprod()
        self.sampler = sampler

    def sample_seconds(self):
        torch, inp, s = self.torch, self.inp, self.sampler
        with torch.no_grad():
            t0 = time.perf_counter()
            ft = s.feature_timesteps(981, self.R, "auto-regressive")[0]
            x = s.add_noise(inp["frame_latents"][0], inp["noise"], torch.full((1,), ft, dtype=torch.long), self.acp)
            taps = self.net(x, ft, inp["prev_cond"][0], None)[1]
            t_feat = time.perf_counter() - t0
            img = {k: torch.cat([v] * self.R, 1) for k, v in taps.items()}
            t0 = time.perf_counter()
            self.net(inp["latents"], 981, inp["text_cond"], img)
            t_den = time.perf_counter() - t0
        return t_feat, t_den

    def describe(self, t_feat, t_den, ddim_steps, n):
        t_step = 3 * (self.R * t_feat + t_den)
        return t_step, (f"{n} x [one feature-pass + one denoise-pass UNet call, 1 CFG lane, R={self.R}: "
                        f"{t_feat:.2f} s + {t_den:.2f} s]; t_step = 3 lanes x (R x feature + denoise) = {t_step:.1f} s; "
                        f"images/s = 1/({ddim_steps} x t_step); fp32 oracle == reference model/*.py, SDPA attention")


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = usable_cores()
    ref = CpuReference(args.frames, threads)
    if args.warmup > 0:
        ref.sample_seconds()
    ts = [ref.sample_seconds() for _ in range(args.steps)]
    t_feat, t_den = sum(t[0] for t in ts) / len(ts), sum(t[1] for t in ts) / len(ts)
    t_step, desc = ref.describe(t_feat, t_den, args.ddim_steps, len(ts))
    value = 1.0 / (args.ddim_steps * t_step)
    line = {
        "impl": "reference", "metric": "images_per_sec_512px_3ref_50step", "value": value, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * args.ddim_steps * t_step * args.batch,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": f"StoryGen denoise loop: SD-1.5 UNet+attn3 (909M), {8 * args.latent}px, {args.frames} prior "
                        f"frames, {args.ddim_steps} DDIM steps, dual CFG (3 lanes), B={args.batch} images/GPU "
                        "(BASELINE cfg2); tokenizer/CLIP/VAE excluded (synthetic latents+embeddings)",
            "global_batch": args.batch * world, "per_gpu_batch": args.batch, "prior_frames": args.frames,
            "ddim_steps": args.ddim_steps, "latent": args.latent, "parallelism": f"dp{world} (batch shard, 1 all-gather)",
            "l2": "working set per step >> 126 MB L2 (no flush needed)"}


def kernel_rooflines(torch, peaks, lanes):
    """Live CUDA-event timing of the dominant kernels in isolation (burst peak as denominator)."""
    from storygen_b200 import ops
    from storygen_b200.ops import BF16, Src, taps3x3
    dev = "cuda"
    out = {}

    def timeit(fn, n_buf, iters=12):
        for i in range(3):
            fn(i % n_buf)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(i % n_buf)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e-3

    # (1) L0 ResBlock conv: 3x3, 320 -> 320 at 64x64 over `lanes` images (the most frequent heavy GEMM);
    #     4 rotating input/output pairs (4 x 2 x 63 MB at 24 lanes) so that L2 does not hold them
    c, h, w = 320, 64, 64
    xs = [torch.randn(lanes, h, w, c, device=dev).to(BF16) for _ in range(4)]
    ys = [torch.empty(lanes, h, w, c, device=dev, dtype=BF16) for _ in range(4)]
    wt = (torch.randn(c, 9 * c, device=dev) * (9 * c) ** -0.5).to(BF16)
    b = torch.zeros(c, device=dev)
    t = timeit(lambda i: ops.gemm([Src.nhwc(xs[i])], taps3x3(0, c), (w, h, lanes), wt, ys[i], bias=b), 4)
    flops = 2.0 * lanes * h * w * c * 9 * c
    # traffic = dram__bytes_read.sum + dram__bytes_write.sum of this kernel inside a step at 24 lanes (ncu --set full,
    # profiles/r01_final_ncu_full_conv.csv): 64.9 MB read + 17.5 MB written < 127.7 MB algorithmic (part of the output is still in L2 at kernel end)
    traffic = 82338816 if lanes == 24 else None
    out["conv3x3_320_L0"] = {"bound": "tensor", "achieved": flops / t / 1e12, "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                             "frac": flops / t / 1e12 / peaks["tf_burst"], "traffic": traffic,
                             "algorithmic_bytes": 2 * lanes * h * w * c * 2 + 9 * c * c * 2,
                             "note": "N = 320 runs as two 160-wide tiles: per 64-deep k-block a CTA pulls A 16 KB + B 20 KB into shared "
                                     "memory for 4 x 80 clk of tcgen05 math = 115 B/clk/SM demanded vs ~80-100 B/clk/SM that TMA delivers "
                                     "from L2 with all SMs active (profiles/r01_ubench_tma_ingest.log): operand-ingest bound, not "
                                     "issue bound (issue streams run at the math rate: profiles/r01_ubench_issue_stream.log)",
                             "flops_per_launch": flops, "ms_per_launch": t * 1e3, "lanes": lanes}
    # (2) L0 image cross-attention core: Nq = 4096, Nk = 3 * 4096, 8 heads x d40
    hd, heads, nq, nk = 40, 8, 4096, 3 * 4096
    ql = max(3, lanes // 2)
    q = torch.randn(ql, nq, c, device=dev).to(BF16)
    k = torch.randn(ql, nk, c, device=dev).to(BF16)
    v = torch.randn(ql, nk, c, device=dev).to(BF16)
    o = torch.empty(ql, nq, c, device=dev, dtype=BF16)
    t = timeit(lambda i: ops.attention(q, k, v, o, ql, heads, hd, nq, nk, (nq * c, c), (nk * c, c), (nk * c, c)), 1, iters=6)
    flops = 4.0 * ql * nq * nk * c
    out["attn3_core_L0"] = {"bound": "tensor", "achieved": flops / t / 1e12, "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                            "frac": flops / t / 1e12 / peaks["tf_burst"], "traffic": None, "flops_per_launch": flops,
                            "ms_per_launch": t * 1e3, "lanes": ql,
                            "note": "d=40: a 128x128 tile needs 16384 exp2 = 1024 clk of MUFU (16/clk/SM; 896 clk with 1 in 8 exps "
                                    "on the FMA pipe) but only ~580 clk of tensor pipe (3 QK^T + 8 PV MMAs): exp-unit bound; ncu: "
                                    "XU pipe 68 %, measured ~1450 SM-clk per tile"}
    del xs, ys, q, k, v, o
    torch.cuda.empty_cache()
    return out


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    from storygen_b200 import DDIMScheduler, StableDiffusionPipeline, UNet2DConditionModel, _lib
    from storygen_b200.distributed import gather_latents, shard_range
    from storygen_b200.synthetic import input_bytes, random_init_, synthetic_inputs, to_device

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    B, R = args.batch, args.frames

    with torch.device("meta"):
        unet = UNet2DConditionModel(cross_attention_dim=768)
    unet = random_init_(unet.to_empty(device=dev), seed=0).eval()
    pipe = StableDiffusionPipeline(None, None, None, unet, DDIMScheduler())
    first, _ = shard_range(world * B, rank, world)
    host_in = synthetic_inputs(B, R, latent_hw=args.latent, first_image=first, pin=True)
    dev_in = to_device(host_in, dev)
    kw = dict(num_inference_steps=args.ddim_steps, guidance_scale=7.0, image_guidance_scale=3.5, stage="auto-regressive")
    host_out = torch.empty(B, 4, args.latent, args.latent).pin_memory()

    def one_step(inputs, to_host):
        lat = pipe.denoise(**inputs, **kw)
        lat_all = gather_latents(lat, world * B)  # the path's only collective (2 MiB at 8 x B=8)
        assert lat_all.shape[0] == world * B
        if to_host:
            host_out.copy_(lat, non_blocking=True)
        return lat

    def timed(inputs, to_host, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            one_step(inputs, to_host)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        one_step(dev_in, False)
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    k0 = pipe.kernel_launches
    ms = timed(dev_in, False, args.steps)
    launches = pipe.kernel_launches - k0
    ms_e2e = timed(host_in, True, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    lat = one_step(dev_in, False)
    finite = bool(torch.isfinite(lat).all().item())

    images = world * B * args.steps
    value = images / (ms * 1e-3)
    e2e_value = images / (ms_e2e * 1e-3)
    if rank == 0:
        roof = kernel_rooflines(torch, peaks, 2 * B * R)
        tf_img = tflop_per_image(R, args.ddim_steps)
        model_tf = tf_img * value / world  # per GPU, reference-equivalent FLOPs (de-duplicated lane NOT credited)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = usable_cores()
            ref = CpuReference(R, threads)
            t_feat, t_den = ref.sample_seconds()
            t_step, desc = ref.describe(t_feat, t_den, args.ddim_steps, 1)
            cpu = {"value": 1.0 / (args.ddim_steps * t_step), "unit": "images/s", "cores": threads, "kind": "port",
                   "sample": desc}
        main = dict(roof["conv3x3_320_L0"])
        line = {
            "metric": "images_per_sec_512px_3ref_50step", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": input_bytes(host_in),
                    "d2h_bytes_per_step": host_out.numel() * 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {k: main[k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic")},
            "roofline_detail": {"peaks": peaks, "kernels": roof,
                                "model": {"tflop_per_image_reference_equivalent": tf_img,
                                          "achieved_tflops_per_gpu": model_tf, "peak_sustained": peaks["tf_sust"],
                                          "frac_of_sustained": model_tf / peaks["tf_sust"]}},
            "cpu_baseline": cpu,
            "outputs_finite": finite,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4, help="images per GPU (cfg2: 4)")
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--ddim-steps", type=int, default=50)
    ap.add_argument("--latent", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1 and args.impl == "ours":
        # convenience: relaunch under torchrun (the driver launches torch.distributed.run itself)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
