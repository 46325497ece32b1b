#!/bin/bash
# ncu --set full of the HBM-bound kernels (K1 fwd/bwd, K4 fwd/bwd, EDT, cull, COM) inside one bench-shaped step.  Under gpurun.
mkdir -p gpurun_out
cat > /tmp/one_step.py <<'PY'
import sys
sys.path.insert(0, ".")
import torch, bench
from doodle_b200 import HelioEnv, CenterOfMass2D
dev = torch.device("cuda:0")
N, R, B = 2000, 256, 4096
helio, targ_pos, targ_norm, area, _ = bench.make_inputs(N, B)
torch.manual_seed(42)
env = HelioEnv(heliostat_pos=helio.to(dev), targ_pos=targ_pos.to(dev), targ_area=area, targ_norm=targ_norm.to(dev), sigma_scale=0.01,
               error_scale_mrad=90.0, resolution=R, batch_size=B, device="cuda:0", check_finite=False, cull=True)
env.reset()
for _ in range(2):
    a = env.noisy_field.initial_action.detach().clone().requires_grad_(True)
    obs, m, mon = env.step(a)
    (m["mse"] + m["dist"] + m["bound"] + m["alignment_loss"]).backward()
x = obs["img"].detach().requires_grad_(True)
CenterOfMass2D()(x).sum().backward()
torch.cuda.synchronize()
print("ok")
PY
python /tmp/one_step.py > gpurun_out/prof_small_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:'geom_fwd|geom_bwd|loss_fwd_kernel|loss_bwd|edt_rows|edt_cols|cull_kernel|com_fwd|com_bwd|image_max' \
    -f -o gpurun_out/prof_small_kernels python /tmp/one_step.py > gpurun_out/prof_small_ncu.log 2>&1
echo "exit $?"; ls -la gpurun_out/prof_small_kernels.ncu-rep
