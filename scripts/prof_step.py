"""A few headline-shaped env steps (N=2000, R=256, B=4096, product defaults) for ncu: the forward and backward splat kernels exactly as
HelioEnv.step launches them (f16x3 operands, fused step).   ncu ... -k regex:splat_bwd_tc -s 2 -c 1 python scripts/prof_step.py"""
import sys
sys.path.insert(0, ".")
import torch, bench
from doodle_b200 import HelioEnv
dev = torch.device("cuda:0")
N, R, B = 2000, 256, int(sys.argv[1]) if len(sys.argv) > 1 else 4096
helio, targ_pos, targ_norm, area, _ = bench.make_inputs(N, B)
torch.manual_seed(42)
env = HelioEnv(heliostat_pos=helio.to(dev), targ_pos=targ_pos.to(dev), targ_area=area, targ_norm=targ_norm.to(dev), sigma_scale=0.01,
               error_scale_mrad=90.0, resolution=R, batch_size=B, device="cuda:0")
env.reset()
for _ in range(4):
    a = env.noisy_field.initial_action.detach().clone().requires_grad_(True)
    obs, m, mon = env.step(a)
    (m["mse"] + m["dist"] + m["bound"] + m["alignment_loss"]).backward()
env.flush_checks()
torch.cuda.synchronize()
print("ok")
