"""Host-side cost breakdown of env.step + backward at the small config (N=50, R=128, B=25): cProfile of 2000 steps."""
import cProfile, pstats, sys, time
sys.path.insert(0, ".")
import torch
from doodle_b200 import HelioEnv

dev = "cuda:0"
torch.manual_seed(0)
N, R, B = 50, 128, 25
helio = torch.rand(N, 3, device=dev) * 10 + 80; helio[:, 2] = 0
env = HelioEnv(helio, torch.tensor([0., -5., 0.], device=dev), (15., 15.), torch.tensor([0., 1., 0.], device=dev),
               sigma_scale=0.1, error_scale_mrad=90.0, resolution=R, batch_size=B, device=dev,
               graph=(False if "--eager" in sys.argv else "auto"))
env.reset()
a0 = env.ideal_normals.flatten(1).clone()
def step(bwd=True):
    a = a0.detach().requires_grad_(True)
    obs, m, mon = env.step(a)
    if bwd:
        (m["mse"] + m["dist"] + m["bound"] + m["alignment_loss"]).backward()
for _ in range(20): step()
torch.cuda.synchronize()
print("graph replay active:", env._step_graph is not None)
def step_one_metric():
    a = a0.detach().requires_grad_(True)
    obs, m, mon = env.step(a)
    m["dist"].backward()
t0 = time.perf_counter()
for _ in range(1000): step_one_metric()
torch.cuda.synchronize()
print(f"single-metric loss: {(time.perf_counter()-t0)*1e3:.1f} us/step wall")
for bwd in (False, True):
    t0 = time.perf_counter()
    for _ in range(1000): step(bwd)
    torch.cuda.synchronize()
    print(f"bwd={bwd}: {(time.perf_counter()-t0)*1e3:.1f} us/step wall")
with torch.no_grad():
    t0 = time.perf_counter()
    for _ in range(1000): env.step(a0)
    torch.cuda.synchronize()
    print(f"no_grad fwd: {(time.perf_counter()-t0)*1e3:.1f} us/step wall")
pr = cProfile.Profile()
pr.enable()
for _ in range(1000): step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(25)
