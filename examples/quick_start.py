#!/usr/bin/env python3
"""The reference README's quick start (README.md:61-95 of l3th4l/DOODLE) on the sm_100a renderer: BASELINE.json configs[0]
shape -- N=50 heliostats, 128x128 receiver, B=25 sun positions, 90 mrad errors, one render + backward.

    python examples/quick_start.py
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from doodle_b200 import HelioField          # or: PYTHONPATH=dropin python -c "from newenv_rl_test_multi_error import HelioField"

dev = "cuda:0"
torch.manual_seed(0)
N, R, B = 50, 128, 25
heliostat_positions = torch.rand(N, 3, device=dev) * 10
heliostat_positions[:, 2] = 0
field = HelioField(heliostat_positions, target_position=torch.tensor([0., -5., 0.], device=dev), target_area=(15., 15.),
                   target_normal=torch.tensor([0., 1., 0.], device=dev), error_scale_mrad=90.0, sigma_scale=0.1,
                   initial_action_noise=0.01, resolution=R, device=dev, max_batch_size=B)
sun = torch.nn.functional.normalize(torch.tensor([[0.5, 0.5, 0.7071]], device=dev) + 0.02 * torch.randn(B, 3, device=dev), dim=1) * 14142.0
ideal = field.calculate_ideal_normals(sun)
field.init_actions(sun)
action = field.initial_action.clone().requires_grad_(True)

def once():
    img, actual = field.render(sun, action, ideal)
    loss = img.pow(2).mean()
    action.grad = None
    loss.backward()
    return img, loss

for _ in range(5):
    once()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    img, loss = once()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 200
print(f"images {tuple(img.shape)}, loss {float(loss):.5f}, |grad| {float(action.grad.abs().max()):.3e}, "
      f"render + backward {dt * 1e6:.0f} us per call ({B * N * R * R / dt / 1e9:.1f} G heliostat-pixel evals/s; "
      f"the reference takes ~0.5 s for the same call on 8 CPU cores, SURVEY.md section 6)")
