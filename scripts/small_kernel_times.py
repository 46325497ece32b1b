"""Per-kernel GPU time at the small configs (composed path, CUDA events around each C-ABI call)."""
import sys
sys.path.insert(0, ".")
import torch
from doodle_b200 import HelioEnv, functional as Fn
dev = "cuda:0"
def run(N, R, B, sigma=0.1):
    torch.manual_seed(0)
    helio = torch.rand(N, 3, device=dev) * 10 + 80; helio[:, 2] = 0
    env = HelioEnv(helio, torch.tensor([0., -5., 0.], device=dev), (15., 15.), torch.tensor([0., 1., 0.], device=dev),
                   sigma_scale=sigma, error_scale_mrad=90.0, resolution=R, batch_size=B, device=dev, fused_step=False, check_finite=False)
    env.reset()
    a0 = env.ideal_normals.flatten(1).clone()
    def step():
        a = a0.detach().requires_grad_(True)
        obs, m, mon = env.step(a)
        (m["mse"] + m["dist"] + m["bound"] + m["alignment_loss"]).backward()
    for _ in range(10): step()
    Fn.reset_profile(True)
    for _ in range(50): step()
    prof = Fn.collect_profile()
    Fn.reset_profile(False)
    print(f"N={N} R={R} B={B}: " + "  ".join(f"{k}={v['avg_ms']*1e3:.1f}us(x{v['n']//50})" for k, v in sorted(prof.items())), flush=True)
for cfg in [(50, 128, 25), (50, 128, 256), (500, 64, 25), (500, 128, 1024), (50, 64, 1024), (5000, 128, 25), (50, 512, 25), (500, 512, 1024//8)]:
    run(*cfg)
