"""A few env.step + backward at one small config, for an ncu launch list (gpu__time_duration per kernel)."""
import sys
sys.path.insert(0, ".")
import torch
from doodle_b200 import HelioEnv
N, R, B = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (50, 128, 25)
dev = "cuda:0"
torch.manual_seed(0)
helio = torch.rand(N, 3, device=dev) * 10 + 80; helio[:, 2] = 0
env = HelioEnv(helio, torch.tensor([0., -5., 0.], device=dev), (15., 15.), torch.tensor([0., 1., 0.], device=dev),
               sigma_scale=0.1, error_scale_mrad=90.0, resolution=R, batch_size=B, device=dev, check_finite=False)
env.reset()
a0 = env.ideal_normals.flatten(1).clone()
for _ in range(4):
    a = a0.detach().requires_grad_(True)
    obs, m, mon = env.step(a)
    (m["mse"] + m["dist"] + m["bound"] + m["alignment_loss"]).backward()
torch.cuda.synchronize()
print("ok")
