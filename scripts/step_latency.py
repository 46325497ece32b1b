"""env.step + backward latency at the small BASELINE configs (C2: N=50,R=128,B=25; C3: B=256) and a few others."""
import sys, time
sys.path.insert(0, ".")
import torch
from doodle_b200 import HelioEnv, functional as Fn

dev = "cuda:0"
def run(N, R, B, sigma=0.1, iters=50, cache=False, fused=True, check=True):
    torch.manual_seed(0)
    helio = torch.rand(N, 3, device=dev) * 10 + 80; helio[:, 2] = 0
    env = HelioEnv(helio, torch.tensor([0., -5., 0.], device=dev), (15., 15.), torch.tensor([0., 1., 0.], device=dev),
                   sigma_scale=sigma, error_scale_mrad=90.0, resolution=R, batch_size=B, device=dev, cache_target=cache, fused_step=fused, check_finite=check)
    env.reset()
    a0 = env.ideal_normals.flatten(1).clone()
    def step():
        a = a0.detach().requires_grad_(True)
        obs, m, mon = env.step(a)
        (m["mse"] + m["dist"] + m["bound"] + m["alignment_loss"]).backward()
        return a.grad
    for _ in range(5): step()
    torch.cuda.synchronize()
    l0 = Fn.launch_count()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): step()
    e1.record(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / iters * 1e3
    gpu = e0.elapsed_time(e1) / iters
    print(f"N={N} R={R} B={B} cache={cache} fused={fused} check_finite={check}: {gpu:.3f} ms/step (wall {wall:.3f}), {B*N*R*R/gpu/1e6:.2f} Geval/s, helio launches/step {(Fn.launch_count()-l0)/iters:.0f}", flush=True)

for cfg in [(50, 128, 25), (50, 128, 256), (500, 64, 25), (500, 128, 1024), (50, 64, 1024), (5000, 128, 25)]:
    run(*cfg)
run(50, 128, 25, cache=True)
run(50, 128, 25, fused=False)
run(50, 128, 25, check=False)
run(50, 128, 25, cache=True, check=False)


def run_graphed(N, R, B, sigma=0.1, iters=200, cache=False):
    from doodle_b200 import GraphedStep
    torch.manual_seed(0)
    helio = torch.rand(N, 3, device=dev) * 10 + 80; helio[:, 2] = 0
    env = HelioEnv(helio, torch.tensor([0., -5., 0.], device=dev), (15., 15.), torch.tensor([0., 1., 0.], device=dev),
                   sigma_scale=sigma, error_scale_mrad=90.0, resolution=R, batch_size=B, device=dev, cache_target=cache)
    env.reset()
    gs = GraphedStep(env)
    a0 = env.ideal_normals.clone()
    for _ in range(5): gs(a0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): gs(a0)
    e1.record(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / iters * 1e3
    gpu = e0.elapsed_time(e1) / iters
    print(f"GRAPHED N={N} R={R} B={B} cache={cache}: {gpu*1e3:.1f} us/step (wall {wall*1e3:.1f} us) = {1e3/gpu:.0f} steps/s, helio kernels/replay {gs.helio_kernels_per_replay}", flush=True)

for cfg in [(50, 128, 25), (50, 128, 256), (500, 64, 25), (50, 64, 1024)]:
    run_graphed(*cfg)
run_graphed(50, 128, 25, cache=True)
