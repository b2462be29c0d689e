"""One eager (no CUDA graph) denoising step of the bench workload between cudaProfilerStart/Stop, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ...
  ncu --profile-from-start off --set full -k regex:<kernel> -c 3 ...
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from storygen_b200 import DDIMScheduler, StableDiffusionPipeline, UNet2DConditionModel
from storygen_b200.pipeline import _DenoiseLoop
from storygen_b200.synthetic import random_init_, synthetic_inputs, to_device

B = int(os.environ.get("SG_BATCH", "4"))
R = int(os.environ.get("SG_FRAMES", "3"))
dev = torch.device("cuda", 0)
with torch.device("meta"):
    unet = UNet2DConditionModel(cross_attention_dim=768)
unet = random_init_(unet.to_empty(device=dev)).eval()
inp = to_device(synthetic_inputs(B, R), dev)
loop = _DenoiseLoop(unet, DDIMScheduler(), B, R, (4, 64, 64), (77, 768), 50, 7.0, 3.5, "auto-regressive")
loop.load(**inp)
loop.par.copy_(loop.tab[0]); loop.par_i.copy_(loop.tab_i[0])
with torch.no_grad():
    loop._step_body()  # warm-up
    torch.cuda.synchronize()
    from storygen_b200 import ops
    ops.TRACE = []
    torch.cuda.profiler.start()
    loop._step_body()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    import json
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(ops.TRACE, open("gpurun_out/trace_step.json", "w"))
print("profiled one denoising step: B=%d R=%d" % (B, R))
