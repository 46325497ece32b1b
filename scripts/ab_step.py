"""Kernel times of a few env steps (product defaults) from the library's CUDA-event profiler, for A/B runs of experimental builds:
    HELIO_LIB_PATH=$PWD/doodle_b200/libhelio_vN.so python scripts/ab_step.py [N R B]"""
import sys
sys.path.insert(0, ".")
import torch, bench
from doodle_b200 import HelioEnv
from doodle_b200 import functional as Fn
dev = torch.device("cuda:0")
N, R, B = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (2000, 256, 4096)
helio, targ_pos, targ_norm, area, _ = bench.make_inputs(N, B)
torch.manual_seed(42)
env = HelioEnv(heliostat_pos=helio.to(dev), targ_pos=targ_pos.to(dev), targ_area=area, targ_norm=targ_norm.to(dev), sigma_scale=0.01,
               error_scale_mrad=90.0, resolution=R, batch_size=B, device="cuda:0", graph=False)
env.reset()
def step():
    a = env.noisy_field.initial_action.detach().clone().requires_grad_(True)
    obs, m, mon = env.step(a)
    (m["mse"] + m["dist"] + m["bound"] + m["alignment_loss"]).backward()
    return a.grad
for _ in range(3):
    g = step()
torch.cuda.synchronize()
Fn.reset_profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(8):
    g = step()
e1.record(); torch.cuda.synchronize()
prof = Fn.collect_profile(); Fn.reset_profile(False)
env.flush_checks()
digest = int(g.contiguous().view(torch.int32).to(torch.int64).sum().item()) & 0xFFFFFFFFFFFF
print(f"N={N} R={R} B={B}: step {e0.elapsed_time(e1) / 8:.3f} ms (profiler on) | " + " ".join(f"{k} {v['avg_ms']:.3f}" for k, v in sorted(prof.items()))
      + f" | grad bits {digest:012x}")
