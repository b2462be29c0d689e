#!/usr/bin/env python
"""Benchmark of StoryGen's denoising hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this framework (sm_100a kernels)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's algorithm on the host CPU cores
  python bench.py --sweep [--gpus N]                       # BASELINE cfg5: R in {1,3,5} x {512,768} px
  python bench.py --batch 8 --gpus 8                       # BASELINE cfg4: global batch 64

Workload (config.workload): SD-1.5-shaped UNet + StoryGen attn3 (909 M params, name-seeded random weights —
no checkpoint exists offline), 512x512 (latent 64x64), 3 prior frames, 50 DDIM steps, dual guidance,
bf16 storage / fp32 accumulation, B images per GPU ("cfg2", B=4).  ONE bench step = one full 50-step
generation of B images on every rank (weak scaling) followed by the single NCCL all-gather of the
final latents; value = images / second over all ranks.  Inputs are synthetic latents + embeddings
(tokenizer / CLIP / VAE are outside the measured path, as in the metric's definition).

Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
One step streams ~1.8 GB of weights and several GB of activations through a 126 MB L2 200 times, so
no explicit L2 flush is needed ("inputs larger than L2").

Roofline: `roofline` describes the kernel with the LARGEST share of a denoising step (profiles/r02_kernel_shares.json,
made from the committed ncu launch list; today attn_kernel<40,2>), timed live with CUDA events in this run;
`roofline_detail.model` is computed from the FLOPs this path EXECUTES (the duplicate CFG lane of the feature passes
and its attn3 K/V projection are not executed and not credited); the reference-equivalent figure is reported
separately as a speed-up factor.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
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
GF_ATTN3_KV_PROJ_PER_FRAME = 25.6  # attn3 to_k/to_v over one prior frame's 16 tapped maps, per K/V lane
GF_UPSAMPLE_CONVS = 67.95  # the 3 Upsample2D convs of one UNet call (2*256*1280*11520 + 2*1024*1280*11520 + 2*4096*640*5760)
UPSAMPLE_EXECUTED_FRACTION = 4.0 / 9.0  # run as four 2x2-tap sub-pixel convolutions of the low-resolution tensor
# dead tail of a feature pass (everything after the last tapped map; the reference computes it and drops it, model/pipeline.py:433-435):
# attn2 to_q + core + to_out, GEGLU FF, proj_out of the last 64x64 transformer block, conv_out
GF_FEATURE_TAIL = (2 * 4096 * 320 * 320 * 3 + 4 * 4096 * 77 * 320 + 2 * 4096 * 320 * 2560 + 2 * 4096 * 1280 * 320
                   + 2 * 4096 * 2880 * 4) / 1e9


def gf_denoise(R: int) -> float:
    return GF_DENOISE_PASS.get(R, 976.9 + (R - 1) * 148.1)


def tflop_per_image_reference(R: int, steps: int) -> float:
    """What the reference executes: 3 CFG lanes x [R feature passes + 1 denoise pass] per step."""
    return 3 * steps * (R * GF_FEATURE_PASS + gf_denoise(R)) / 1e3


def tflop_per_image_executed(R: int, steps: int) -> float:
    """What THIS path executes per image: feature passes on 2 of the 3 CFG lanes (the third duplicates the second,
    model/pipeline.py:429-430) and only up to their last tapped map, 3 denoise lanes of which two share one attn3 K/V
    projection (kv_map), and the up-sampling convolutions as sub-pixel phases (4/9 of their dense FLOPs)."""
    feature_lanes, denoise_lanes, kv_lanes = 2, 3, 2
    up_saved = GF_UPSAMPLE_CONVS * (1.0 - UPSAMPLE_EXECUTED_FRACTION)  # per lane and UNet call
    per_step = feature_lanes * R * (GF_FEATURE_PASS - up_saved - GF_FEATURE_TAIL) + denoise_lanes * (gf_denoise(R) - up_saved) \
        - (denoise_lanes - kv_lanes) * R * GF_ATTN3_KV_PROJ_PER_FRAME
    return steps * per_step / 1e3


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


def load_json(rel):
    p = os.path.join(ROOT, rel)
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return None


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

    The unit of work is ONE UNet call of a full denoising step of ONE image (B = 1): the 3 CFG lanes batched into
    one call, exactly as model/pipeline.py:418-453 issues them — R feature-pass calls (one per prior frame, each
    noised to its own timestep) and 1 denoise-pass call conditioned on the R x 16 tapped feature maps.  A bench
    "step" of the reference arm is one such call, cycling  denoise, feature_0 .. feature_{R-1}; R + 1 consecutive
    steps are one complete denoising step (11.05 TFLOP at R = 3).  images/s = 1 / (ddim_steps x t_step) with
    t_step = R x median(feature call) + median(denoise call).
    Attention: torch SDPA (same math as the reference's default baddbmm -> softmax -> bmm path, without
    materialising 4.8 GB of scores per L0 attn3 call)."""

    def __init__(self, R: int, threads: int, latent: int = 64):
        import torch

        from oracle import sampler
        from oracle.unet import SD15_CONFIG, OracleUNet
        from oracle.weights import load_seeded
        self.torch, self.R, self.sampler = torch, R, sampler
        torch.set_num_threads(threads)
        torch.set_grad_enabled(False)
        self.net = load_seeded(OracleUNet(**SD15_CONFIG).eval(), 0)
        self.inp = sampler.synthetic_inputs(1, R, latent_hw=latent)
        self.acp = sampler.alphas_cumprod()
        self.t = sampler.ddim_timesteps(50)[0]
        self.taps = None

    def feature_call(self, i: int) -> float:
        torch, inp, s = self.torch, self.inp, self.sampler
        ft = s.feature_timesteps(self.t, self.R, "auto-regressive")[i]
        t0 = time.perf_counter()
        tt = torch.full((1,), ft, dtype=torch.long)
        noisy_frame = s.add_noise(inp["frame_latents"][i], inp["noise"], tt, self.acp)
        noisy_zero = s.add_noise(inp["zero_latents"], inp["noise"], tt, self.acp)
        x = torch.cat([noisy_zero, noisy_frame, noisy_frame])
        txt = torch.cat([inp["prev_uncond"][i], inp["prev_cond"][i], inp["prev_cond"][i]])
        taps = self.net(x, ft, txt, None)[1]
        dt = time.perf_counter() - t0
        if self.taps is None:
            self.taps = taps
        return dt

    def denoise_call(self) -> float:
        torch, inp = self.torch, self.inp
        if self.taps is None:
            self.feature_call(0)
        img = {k: torch.cat([v] * self.R, 1) for k, v in self.taps.items()}  # shapes of the real step; values irrelevant
        t0 = time.perf_counter()
        txt = torch.cat([inp["text_uncond"], inp["text_uncond"], inp["text_cond"]])
        self.net(torch.cat([inp["latents"]] * 3), self.t, txt, img)
        return time.perf_counter() - t0

    def call(self, k: int):
        """k-th call of the cycle: denoise, feature_0, ..., feature_{R-1}."""
        i = k % (self.R + 1)
        return ("denoise", self.denoise_call()) if i == 0 else ("feature", self.feature_call(i - 1))

    def describe(self, tf, td, ddim_steps):
        t_step = self.R * statistics.median(tf) + statistics.median(td)
        return t_step, (f"B=1 denoising step as the reference issues it: {len(tf)} feature-pass + {len(td)} denoise-pass UNet "
                        f"calls at 3 CFG lanes (R={self.R}); median {statistics.median(tf):.2f} s / {statistics.median(td):.2f} s per call; "
                        f"t_step = R x feature + denoise = {t_step:.1f} s; images/s = 1/({ddim_steps} x t_step); "
                        "fp32 oracle == reference model/*.py, SDPA attention")


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = usable_cores()
    ref = CpuReference(args.frames, threads, args.latent)
    for k in range(args.warmup):
        ref.call(k)
    tf, td, wall = [], [], 0.0
    for k in range(args.steps):
        kind, dt = ref.call(k)
        (td if kind == "denoise" else tf).append(dt)
        wall += dt
    extra = []
    if not tf:  # K too small to contain both kinds: one extra sample, reported, outside the K timed steps
        tf.append(ref.feature_call(0))
        extra.append("feature")
    if not td:
        td.append(ref.denoise_call())
        extra.append("denoise")
    t_step, desc = ref.describe(tf, td, args.ddim_steps)
    if extra:
        desc += f"; extra untimed-in-K samples: {extra}"
    value = 1.0 / (args.ddim_steps * t_step)
    line = {
        "impl": "reference", "metric": "images_per_sec_512px_3ref_50step", "value": value, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "step_definition": "one UNet call (3 CFG lanes) of a B=1 denoising step, cycling denoise, feature_0..feature_{R-1}",
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    cfg = "cfg2" if args.batch == 4 else ("cfg4" if args.batch == 8 else f"B={args.batch}")
    return {"workload": f"StoryGen denoise loop: SD-1.5 UNet+attn3 (909M), {8 * args.latent}px, {args.frames} prior "
                        f"frames, {args.ddim_steps} DDIM steps, dual CFG (3 lanes), B={args.batch} images/GPU "
                        f"(BASELINE {cfg}); tokenizer/CLIP/VAE excluded (synthetic latents+embeddings)",
            "global_batch": args.batch * world, "per_gpu_batch": args.batch, "prior_frames": args.frames,
            "ddim_steps": args.ddim_steps, "latent": args.latent, "parallelism": f"dp{world} (batch shard, 1 all-gather)",
            "l2": "working set per step >> 126 MB L2 (no flush needed)"}


# ------------------------------------------------------------------------------------------------------------------
# live kernel timings (CUDA events on torch's current stream = the stream the kernels are launched on)
# ------------------------------------------------------------------------------------------------------------------
def _timeit(torch, fn, n_buf, iters=10):
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


def time_attention(torch, lanes, heads, hd, nq, nk, iters=6):
    """One attention-core launch: (seconds, flops, algorithmic bytes)."""
    from storygen_b200 import ops
    from storygen_b200.ops import BF16
    c = heads * hd
    q = torch.randn(lanes, nq, c, device="cuda").to(BF16)
    k = torch.randn(lanes, nk, c, device="cuda").to(BF16)
    v = torch.randn(lanes, nk, c, device="cuda").to(BF16)
    o = torch.empty(lanes, nq, c, device="cuda", dtype=BF16)
    t = _timeit(torch, lambda i: ops.attention(q, k, v, o, lanes, heads, hd, nq, nk, (nq * c, c), (nk * c, c), (nk * c, c)), 1, iters)
    return t, 4.0 * lanes * nq * nk * c, (2 * nq + 2 * nk) * c * 2.0 * lanes


def kernel_rooflines(torch, peaks, B, R, latent, ms_per_denoising_step):
    """Live timing of the dominant kernels in isolation (burst peak as denominator).  Shares are isolated-kernel time x
    launches per denoising step / measured step time; the ncu launch list under profiles/ is the authoritative split."""
    from storygen_b200 import ops
    from storygen_b200.ops import BF16, Src, taps3x3
    dev = "cuda"
    out = {}
    hw = latent * latent
    nf, nd = 2 * B * R, 3 * B

    # ---- attn_kernel<40,2>: the L0 attention cores of a step (5 transformer blocks at L0 per UNet call)
    shapes = [("attn3 image cross-attention, denoise pass", nd, hw, R * hw, 5),
              ("attn1 self-attention, feature pass", nf, hw, hw, 5),
              ("attn1 self-attention, denoise pass", nd, hw, hw, 5)]
    t_sum = f_sum = b_sum = 0.0
    per_shape = []
    for name, lanes, nq, nk, count in shapes:
        t, fl, by = time_attention(torch, lanes, 8, 40, nq, nk)
        per_shape.append({"shape": name, "lanes": lanes, "nq": nq, "nk": nk, "launches_per_step": count,
                          "ms_per_launch": t * 1e3, "tflops": fl / t / 1e12, "algorithmic_gbs": by / t / 1e9})
        t_sum += count * t
        f_sum += count * fl
        b_sum += count * by
    ncu = load_json("profiles/r02_ncu_attn40.json") or {}
    out["attn_kernel<40,2>"] = {
        "bound": "tensor", "achieved": f_sum / t_sum / 1e12, "peak": peaks["tf_burst"], "unit": "TFLOP/s",
        "frac": f_sum / t_sum / 1e12 / peaks["tf_burst"], "traffic": ncu.get("dram_bytes_per_launch"),
        "share_of_step": t_sum * 1e3 / ms_per_denoising_step, "launches_per_step": 15,
        "attn_tc_util_pct": 100.0 * f_sum / t_sum / 1e12 / peaks["tf_burst"],
        "tensor_pipe_pct": ncu.get("tensor_pipe_pct"), "xu_pipe_pct": ncu.get("xu_pipe_pct"),
        "ncu_source": ncu.get("source"), "algorithmic_gbs": b_sum / t_sum / 1e9, "shapes": per_shape,
        "note": "d=40 attention is exp-unit bound, not tensor bound: a 128x128 score tile needs 16384 exp2 (MUFU 16/clk/SM "
                "= 1024 clk) against ~580 clk of tcgen05 math; achieved/peak is therefore reported against the bf16 tensor "
                "peak as the metric asks (attn TC-util), with the XU / tensor pipe percentages of the ncu capture beside it"}

    # ---- L0 ResBlock conv 3x3 320 -> 320 over the feature-pass lanes (most frequent heavy GEMM, gemm_kernel<160,0>)
    c, h, w, lanes = 320, latent, latent, nf
    xs = [torch.randn(lanes, h, w, c, device=dev).to(BF16) for _ in range(4)]
    ys = [torch.empty(lanes, h, w, c, device=dev, dtype=BF16) for _ in range(4)]
    wt = (torch.randn(c, 9 * c, device=dev) * (9 * c) ** -0.5).to(BF16)
    b = torch.zeros(c, device=dev)
    t = _timeit(torch, lambda i: ops.gemm([Src.nhwc(xs[i])], taps3x3(0, c), (w, h, lanes), wt, ys[i], bias=b), 4, 12)
    flops = 2.0 * lanes * h * w * c * 9 * c
    ncu_c = load_json("profiles/r02_ncu_conv320.json") or {}
    out["conv3x3_320_L0"] = {"bound": "tensor", "achieved": flops / t / 1e12, "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                             "frac": flops / t / 1e12 / peaks["tf_burst"], "traffic": ncu_c.get("dram_bytes_per_launch"),
                             "algorithmic_bytes": 2 * lanes * h * w * c * 2 + 9 * c * c * 2,
                             "tensor_pipe_pct": ncu_c.get("tensor_pipe_pct"), "flops_per_launch": flops,
                             "ms_per_launch": t * 1e3, "lanes": lanes}
    del xs, ys
    torch.cuda.empty_cache()
    return out


def eager_unet_latency(torch, unet, B, R, latent, dev):
    """Latency of the drop-in `unet(...)` call WITHOUT the pipeline's CUDA graph (what a caller of the class surface
    gets): the feature pass over 2BR lanes and the denoise pass over 3B lanes of a step."""
    g = torch.Generator("cpu").manual_seed(9)
    nf, nd = 2 * B * R, 3 * B
    xf = torch.randn(nf, 4, latent, latent, generator=g).to(dev)
    tf = torch.randn(nf, 77, 768, generator=g).to(dev)
    xd, td = xf[:nd].contiguous(), tf[:nd].contiguous()

    def once(fn):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3

    taps = unet(xf, 500, tf, return_dict=False)[1]
    cond = {k: v.view(2 * B, -1, v.shape[-1])[B:].repeat(3, 1, 1)[:nd].contiguous() for k, v in taps.items()}
    ms_f = once(lambda: unet(xf, 500, tf, return_dict=False))
    ms_d = once(lambda: unet(xd, 500, td, image_hidden_states=cond, return_dict=False))
    return {"feature_pass_ms": ms_f, "feature_lanes": nf, "denoise_pass_ms": ms_d, "denoise_lanes": nd,
            "note": "host-synchronous wall time of one eager call (tensor maps re-encoded per launch, caching allocator)"}


def build_model(torch, dev):
    from storygen_b200 import UNet2DConditionModel
    from storygen_b200.synthetic import seeded_init_
    with torch.device("meta"):
        unet = UNet2DConditionModel(cross_attention_dim=768)
    # name-seeded CPU weights (identical on every rank and in the fp32 oracle that made tests/golden/bench_steps_fp32.pt)
    return seeded_init_(unet.to_empty(device="cpu"), seed=0).to(dev).eval()


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    from storygen_b200 import DDIMScheduler, StableDiffusionPipeline
    from storygen_b200.distributed import gather_latents, shard_range
    from storygen_b200.synthetic import input_bytes, synthetic_inputs, to_device

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    unet = build_model(torch, dev)
    pipe = StableDiffusionPipeline(None, None, None, unet, DDIMScheduler())

    def measure(B, R, latent, steps, warmup, with_e2e=True, sampler=None):
        first, _ = shard_range(world * B, rank, world)
        host_in = synthetic_inputs(B, R, latent_hw=latent, first_image=first, pin=True)
        dev_in = to_device(host_in, dev)
        kw = dict(num_inference_steps=args.ddim_steps, guidance_scale=7.0, image_guidance_scale=3.5, stage="auto-regressive")
        host_out = torch.empty(B, 4, latent, latent).pin_memory()

        def one_step(inputs, to_host, callback=None):
            lat = pipe.denoise(**inputs, **kw, callback=callback)
            lat_all = gather_latents(lat, world * B)  # the path's only collective (2 MiB at 8 x B=8)
            assert lat_all.shape[0] == world * B
            if to_host:
                host_out.copy_(lat, non_blocking=True)
            return lat

        def timed(inputs, to_host, n):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                one_step(inputs, to_host)
            e1.record()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return float(ms.item())

        for _ in range(warmup):
            one_step(dev_in, False)
        torch.cuda.synchronize()
        if sampler is not None:
            sampler.start()
        k0 = pipe.kernel_launches
        ms = timed(dev_in, False, steps)
        launches = pipe.kernel_launches - k0
        ms_e2e = timed(host_in, True, steps) if with_e2e else None
        return dict(ms=ms, ms_e2e=ms_e2e, launches=launches, one_step=one_step, dev_in=dev_in, host_in=host_in,
                    host_out=host_out, first=first)

    if args.sweep:
        run_sweep(args, torch, dist, measure, peaks, rank, world)
        if world > 1:
            dist.destroy_process_group()
        return

    B, R = args.batch, args.frames
    sampler = ClockSampler(local_rank) if rank == 0 else None
    m = measure(B, R, args.latent, args.steps, args.warmup, True, sampler)
    clocks = sampler.stop() if rank == 0 else None
    ms, ms_e2e = m["ms"], m["ms_e2e"]

    # parity of THIS run's outputs: latents after DDIM steps 1 and 2 against the stored fp32 oracle trajectory
    caught = []
    lat = m["one_step"](m["dev_in"], False, callback=lambda s, t, l: caught.append(l.clone()) if s < 2 else None)
    finite = bool(torch.isfinite(lat).all().item())
    parity = {"checked": False}
    gold_path = os.path.join(ROOT, "tests", "golden", "bench_steps_fp32.pt")
    if rank == 0 and os.path.exists(gold_path):
        g = torch.load(gold_path)
        if g["frames"] == R and g["latent"] == args.latent and g["ddim_steps"] == args.ddim_steps and g["images"] >= B \
                and m["first"] == 0:
            parity = {"checked": True, "fixture": "tests/golden/bench_steps_fp32.pt (fp32 oracle, oracle/make_bench_golden.py)"}
            for s, (got, ref) in enumerate(zip(caught, g["latents_step"])):
                ref = ref[:B].to(dev)
                parity[f"max_abs_delta_step{s + 1}"] = float((got - ref).abs().max().item())
                parity[f"ref_std_step{s + 1}"] = float(ref.std().item())

    images = world * B * args.steps
    value = images / (ms * 1e-3)
    e2e_value = images / (ms_e2e * 1e-3)
    if rank == 0:
        from storygen_b200.synthetic import input_bytes as _ib
        ms_denoising_step = ms / args.steps / args.ddim_steps
        roof = kernel_rooflines(torch, peaks, B, R, args.latent, ms_denoising_step)
        shares = load_json("profiles/r02_kernel_shares.json") or {}
        dominant = shares.get("dominant", "attn_kernel<40,2>")
        main = dict(roof.get(dominant, roof["attn_kernel<40,2>"]))
        if args.latent != 64:
            raise SystemExit("the FLOP model (SURVEY.md Appendix B) is tabulated for 512 px; use --sweep for 768 px")
        tf_exec = tflop_per_image_executed(R, args.ddim_steps)
        tf_ref = tflop_per_image_reference(R, args.ddim_steps)
        model_tf = tf_exec * value / world  # per GPU, EXECUTED FLOPs
        eager = eager_unet_latency(torch, unet, B, R, args.latent, dev)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = usable_cores()
            ref = CpuReference(R, threads, args.latent)
            tf_, td_ = [ref.feature_call(0)], [ref.denoise_call()]
            t_step, desc = ref.describe(tf_, td_, args.ddim_steps)
            cpu = {"value": 1.0 / (args.ddim_steps * t_step), "unit": "images/s", "cores": threads, "kind": "port",
                   "sample": desc}
        line = {
            "metric": "images_per_sec_512px_3ref_50step", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": _ib(m["host_in"]),
                    "d2h_bytes_per_step": m["host_out"].numel() * 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": m["launches"],
            "clocks": clocks,
            "roofline": {"kernel": dominant if dominant in roof else "attn_kernel<40,2>",
                         **{k: main.get(k) for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "share_of_step",
                                                    "attn_tc_util_pct", "tensor_pipe_pct", "xu_pipe_pct")}},
            "attn_tc_util_pct": roof["attn_kernel<40,2>"]["attn_tc_util_pct"],
            "roofline_detail": {"peaks": peaks, "kernels": roof, "kernel_shares": shares.get("kernels"),
                                "model": {"tflop_per_image_executed": tf_exec,
                                          "achieved_tflops_per_gpu": model_tf, "peak_sustained": peaks["tf_sust"],
                                          "frac_of_sustained": model_tf / peaks["tf_sust"],
                                          "frac_of_burst": model_tf / peaks["tf_burst"],
                                          "tflop_per_image_reference_equivalent": tf_ref,
                                          "work_removed_speedup": tf_ref / tf_exec,
                                          "note": "executed = 2R feature lanes (each up to its last tapped map; the reference drops the rest "
                                                  "of that call's output) + 3 denoise lanes - 1 shared attn3 K/V projection per "
                                                  "image-step, up-sampling convs at 4/9 (sub-pixel phases); the reference executes "
                                                  "3R + 3 full lanes"},
                                "ms_per_denoising_step_graph": ms_denoising_step, "eager_unet": eager},
            "cpu_baseline": cpu,
            "parity": parity,
            "outputs_finite": finite,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_sweep(args, torch, dist, measure, peaks, rank, world):
    """BASELINE cfg5: prior-frame count {1,3,5} x resolution {512,768}: images/s plus the L0 image cross-attention
    core's achieved TFLOP/s (vs the tensor roofline) and algorithmic HBM GB/s, per point."""
    points = []
    B = args.batch
    for latent in (64, 96):
        for R in (1, 3, 5):
            m = measure(B, R, latent, 1, 1, with_e2e=False)
            row = {"prior_frames": R, "px": 8 * latent, "per_gpu_batch": B,
                   "images_per_sec": world * B / (m["ms"] * 1e-3), "ms_per_generation": m["ms"], "gpu_launches": m["launches"]}
            if rank == 0:
                hw = latent * latent
                t, fl, by = time_attention(torch, 3 * B, 8, 40, hw, R * hw, iters=3)
                row["attn3_L0_core"] = {"lanes": 3 * B, "nq": hw, "nk": R * hw, "ms": t * 1e3, "tflops": fl / t / 1e12,
                                        "frac_of_bf16_peak": fl / t / 1e12 / peaks["tf_burst"],
                                        "algorithmic_hbm_gbs": by / t / 1e9, "hbm_frac": by / t / 1e9 / peaks["hbm"]}
                tf_exec = tflop_per_image_executed(R, args.ddim_steps)
                if latent == 64:
                    row["model_tflops_executed_per_gpu"] = tf_exec * row["images_per_sec"] / world
            points.append(row)
            torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps({"metric": "sweep_images_per_sec", "unit": "images/s", "n_gpus": world, "dtype": "bf16",
                          "data": "synthetic", "config": {"workload": "BASELINE cfg5 sweep: R in {1,3,5} x {512,768} px, "
                                                          f"{args.ddim_steps} DDIM steps, B={B}/GPU", "peaks": peaks},
                          "sweep": points}), flush=True)


def run_train(args, rank, local_rank, world):
    """BASELINE cfg3: one stage-2 training step (train_StorySalon_stage2.py:258-332) at 512 px, R prior frames, batch b, bf16
    autocast: R feature passes (inference engine: no trainable parameter on that path), the denoise pass in train() mode
    (attn3 core forward + backward on the tcgen05 kernels, frozen layers as torch ops for their dgrad), masked-MSE loss,
    backward, AdamW step on the 64 attn3 matrices.  N > 1: data parallel, batch b per rank, the 198 MB of attn3 gradients
    averaged with one flat all-reduce per step (storygen_b200.distributed.allreduce_gradients); time = max over ranks."""
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F

    from storygen_b200.distributed import allreduce_gradients, shard_range
    from storygen_b200.synthetic import synthetic_inputs, to_device
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    unet = build_model(torch, dev)
    unet.requires_grad_(False)
    trainable = []
    for name, mod in unet.named_modules():
        if name.endswith("attn3"):
            for p in mod.parameters():
                p.requires_grad = True
                trainable.append(p)
    opt = torch.optim.AdamW(trainable, lr=1e-5, fused=True)
    unet.train()
    b, R = args.batch, args.frames
    inp = to_device(synthetic_inputs(b, R, latent_hw=args.latent, first_image=shard_range(world * b, rank, world)[0]), dev)
    from storygen_b200.scheduler import DDIMScheduler
    acp = DDIMScheduler().alphas_cumprod.to(dev)
    g = torch.Generator(device=dev).manual_seed(rank)
    grad_bytes = 0

    def step():
        t = torch.randint(0, 1000, (b,), device=dev, generator=g)
        rt = (t / 10).long()
        noise = torch.randn(inp["latents"].shape, device=dev, generator=g)
        a = acp[t].view(-1, 1, 1, 1)
        noisy = a.sqrt() * inp["latents"] + (1 - a).sqrt() * noise
        feats = []
        for i in range(R):
            ti = rt * (R - i)
            ai = acp[ti].view(-1, 1, 1, 1)
            ref = ai.sqrt() * inp["frame_latents"][i] + (1 - ai).sqrt() * inp["noise"]
            feats.append(unet(ref, ti, inp["prev_cond"][i], return_dict=False)[1])
        cond = {k: torch.cat([f[k] for f in feats], dim=1) for k in feats[0]}
        pred = unet(noisy, t, inp["text_cond"], image_hidden_states=cond, return_dict=False)[0]
        loss = F.mse_loss(pred.float(), noise)
        loss.backward()
        nonlocal grad_bytes
        grad_bytes = allreduce_gradients(trainable)
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    for _ in range(max(1, args.warmup)):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    print(json.dumps({"metric": "train_images_per_sec_512px_stage2", "value": world * b / (ms * 1e-3), "unit": "images/s",
                      "n_gpus": world, "scaling": "weak", "grad_allreduce_bytes_per_step": grad_bytes,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "dtype": "bf16",
                      "data": "synthetic", "loss": float(loss.item()),
                      "config": {"workload": f"StoryGen stage-2 training step (BASELINE cfg3): SD-1.5 UNet+attn3, 512px, {R} prior "
                                             f"frames, batch {b}, bf16 autocast, fwd+bwd through the context attention, AdamW on attn3",
                                 "custom_kernels": "attn3 attention core forward + backward (sg_attention / sg_attention_bwd; head dims "
                                                   "40 and 80; the two d=160 levels via torch SDPA), feature passes on the inference engine",
                                 "scaffolding": "frozen layers as torch ops under autocast (their dgrad via torch autograd)",
                                 "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4, help="images per GPU (cfg2: 4; cfg4: 8)")
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--ddim-steps", type=int, default=50)
    ap.add_argument("--latent", type=int, default=64)
    ap.add_argument("--sweep", action="store_true", help="BASELINE cfg5: R in {1,3,5} x {512,768} px")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--train", action="store_true", help="BASELINE cfg3: stage-2 training step (use --batch 8)")
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
    elif args.train:
        run_train(args, rank, local_rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
