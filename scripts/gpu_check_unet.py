"""Bring-up check (run under gpurun): CUDA UNet vs the fp32 oracle on the same seeded weights/inputs."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.unet import OracleUNet, TINY_CONFIG, SD15_CONFIG
from oracle.weights import load_seeded
from storygen_b200 import UNet2DConditionModel
from storygen_b200 import _lib

def run(name, cfg, n, R, hw=64, seed=0):
    torch.manual_seed(1)
    ora = load_seeded(OracleUNet(**cfg).eval(), seed)
    mod = UNet2DConditionModel(**cfg)
    mod.load_state_dict(ora.state_dict())
    mod = mod.cuda().eval()
    xd = cfg["cross_attention_dim"]
    x = torch.randn(n, 4, hw, hw); t = torch.tensor([98.0, 294.0, 981.0][:n] + [5.0] * max(0, n - 3)); txt = torch.randn(n, 77, xd)
    with torch.no_grad():
        t0 = time.time(); o_s, o_taps = ora(x, t.long(), txt); t_or = time.time() - t0
        torch.cuda.synchronize(); t0 = time.time()
        c_s, c_taps = mod(x.cuda(), t.cuda(), txt.cuda(), return_dict=False)
        torch.cuda.synchronize(); t_cu = time.time() - t0
    err = (c_s.cpu() - o_s).abs().max().item()
    terr = max((c_taps[k].float().cpu() - o_taps[k]).abs().max().item() for k in o_taps)
    print(f"[{name}] feature pass: max|d sample|={err:.4g} (std {o_s.std():.3g}) max|d taps|={terr:.4g} "
          f"(std {max(v.std().item() for v in o_taps.values()):.3g}) oracle {t_or:.1f}s cuda {t_cu*1e3:.0f}ms", flush=True)
    # denoise pass: condition on R noisy copies of the oracle taps (identical inputs for both)
    img = {k: torch.cat([v * (1 + 0.1 * r) for r in range(R)], 1) for k, v in o_taps.items()}
    with torch.no_grad():
        o2 = ora(x, 981, txt, img)[0]
        c2, d2 = mod(x.cuda(), 981, txt.cuda(), image_hidden_states={k: v.cuda() for k, v in img.items()}, return_dict=False)
    err2 = (c2.cpu() - o2).abs().max().item()
    print(f"[{name}] denoise pass R={R}: max|d sample|={err2:.4g} (std {o2.std():.3g}) dict={len(d2)} launches={_lib.launch_count()}", flush=True)

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    if which in ("tiny", "all"):
        run("tiny n=2", TINY_CONFIG, 2, 2)
        run("tiny n=3 hw=32", TINY_CONFIG, 3, 1, hw=32)
    if which in ("full", "all"):
        run("sd15 n=1", SD15_CONFIG, 1, 3)
