#!/usr/bin/env python3
"""NCCL parity of the sharded env (run under torchrun with W ranks): the W rank-local slices must reproduce the
single-process HelioEnv built from the same seed -- same suns, errors, images, global metrics and action gradients.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_parity.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from doodle_b200 import HelioEnv
from doodle_b200.dist import make_sharded_env, shard_bounds

rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
dist.init_process_group("nccl", device_id=dev)
ok = True
for (N, R, B, mask) in [(40, 64, 8 * world, False), (300, 256, 4 * world, False), (40, 64, 8 * world, True)]:
    g = torch.Generator().manual_seed(5)
    helio = torch.rand(N, 3, generator=g) * 10 + 80; helio[:, 2] = 0
    kw = dict(heliostat_pos=helio.to(dev), targ_pos=torch.tensor([0., -5., 0.], device=dev), targ_area=(15., 15.),
              targ_norm=torch.tensor([0., 1., 0.], device=dev), sigma_scale=0.05, error_scale_mrad=90.0, resolution=R, device=str(dev),
              use_error_mask=mask)
    env = make_sharded_env(HelioEnv, global_batch_size=B, seed=123, **kw)
    torch.manual_seed(123)
    full = HelioEnv(batch_size=B, **kw)                       # every rank also builds the global env (same seed)
    lo, hi = shard_bounds(B, rank, world)
    assert torch.equal(env.sun_pos, full.sun_pos[lo:hi])
    # nothing is copied in: the sharded env draws the GLOBAL error / init-noise tensors and keeps its rows, all-reduces
    # ref_min / ref_max, and builds its distance maps from its own target images
    assert torch.equal(env.noisy_field.batch_error_angles_mrad, full.noisy_field.batch_error_angles_mrad[lo:hi]), "constructor errors"
    assert torch.equal(env.distance_maps, full.distance_maps[lo:hi]), "distance maps"
    assert torch.equal(env.ref_min, full.ref_min) and torch.equal(env.ref_max, full.ref_max), "ref_min / ref_max"
    torch.manual_seed(77); o_loc = env.reset()
    torch.manual_seed(77); o_full = full.reset()
    assert torch.equal(env.noisy_field.batch_error_angles_mrad, full.noisy_field.batch_error_angles_mrad[lo:hi]), "reset errors"
    assert torch.equal(env.noisy_field.initial_action, full.noisy_field.initial_action[lo:hi]), "init_actions noise"
    assert torch.equal(o_loc["img"], o_full["img"][lo:hi]), "reset image"
    torch.manual_seed(7)
    act_full = torch.nn.functional.normalize(full.ref_field.initial_action.view(B, N, 3) + 0.02 * torch.randn(B, N, 3, device=dev), dim=2)
    a_full = act_full.clone().requires_grad_(True)
    a_loc = act_full[lo:hi].clone().requires_grad_(True)
    of, mf, _ = full.step(a_full)
    ol, ml, _ = env.step(a_loc)
    w = dict(mse=1.0, dist=0.01, bound=1.0, alignment_loss=1.0)
    gf, = torch.autograd.grad(sum(w[k] * mf[k] for k in w), a_full)
    gl, = torch.autograd.grad(sum(w[k] * ml[k] for k in w), a_loc)
    img_ok = torch.equal(ol["img"], of["img"][lo:hi])
    met = {k: (float(ml[k]), float(mf[k])) for k in w}
    met_ok = all(abs(a - b) <= 2e-5 * abs(b) + 1e-9 for a, b in met.values())
    gerr = float((gl - gf[lo:hi]).abs().max() / gf.abs().max())
    good = img_ok and met_ok and gerr < 1e-5
    flag = torch.tensor([1.0 if good else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"N={N} R={R} B={B} mask={mask} world={world}: images equal {img_ok}, metrics {met_ok}, grad rel err {gerr:.2e} -> "
              f"{'ok' if flag.item() == 1 else 'MISMATCH'}", flush=True)
    ok = ok and flag.item() == 1
# transparent graph replay on the sharded env: the packed all-reduce is captured into the forward graph (NCCL); several steps,
# two of them alive at once, against the eager single-process env
N, R, B = 50, 128, 8 * world
g = torch.Generator().manual_seed(6)
helio = torch.rand(N, 3, generator=g) * 10 + 80; helio[:, 2] = 0
kw = dict(heliostat_pos=helio.to(dev), targ_pos=torch.tensor([0., -5., 0.], device=dev), targ_area=(15., 15.),
          targ_norm=torch.tensor([0., 1., 0.], device=dev), sigma_scale=0.05, error_scale_mrad=90.0, resolution=R, device=str(dev))
env = make_sharded_env(HelioEnv, global_batch_size=B, seed=321, graph=True, **kw)
torch.manual_seed(321)
full = HelioEnv(batch_size=B, graph=False, **kw)
lo, hi = shard_bounds(B, rank, world)
good = True
for rep in range(3):
    torch.manual_seed(50 + rep); env.reset()
    torch.manual_seed(50 + rep); full.reset()
    torch.manual_seed(9 + rep)
    acts = [torch.nn.functional.normalize(full.ideal_normals + 0.02 * torch.randn(B, N, 3, device=dev), dim=2) for _ in range(2)]
    la, lf, leaves_a, leaves_f = [], [], [], []
    for a in acts:
        al = a[lo:hi].clone().requires_grad_(True); af = a.clone().requires_grad_(True)
        _, ml, _ = env.step(al); _, mf, _ = full.step(af)
        la.append(ml["mse"] + 0.01 * ml["dist"] + ml["bound"] + ml["alignment_loss"]); lf.append(mf["mse"] + 0.01 * mf["dist"] + mf["bound"] + mf["alignment_loss"])
        leaves_a.append(al); leaves_f.append(af)
    sum(la).backward(); sum(lf).backward()
    for k in range(2):
        met_ok = abs(float(la[k]) - float(lf[k])) <= 2e-5 * abs(float(lf[k]))
        gerr = float((leaves_a[k].grad - leaves_f[k].grad[lo:hi]).abs().max() / leaves_f[k].grad.abs().max())
        good = good and met_ok and gerr < 1e-5
replayed = env._step_graph is not None
flag = torch.tensor([1.0 if (good and replayed) else 0.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"sharded graph replay N={N} R={R} B={B} world={world}: replayed {replayed}, parity {good} -> {'ok' if flag.item() == 1 else 'MISMATCH'}", flush=True)
ok = ok and flag.item() == 1
env.close()                       # graphs that captured NCCL kernels must go before the communicator does
del env, full
import gc, threading
gc.collect()
torch.cuda.synchronize()
code = 0 if ok else 1
t = threading.Timer(30.0, lambda: os._exit(code)); t.daemon = True; t.start()      # never let a teardown hang cost GPU time
dist.destroy_process_group()
sys.stdout.flush()
os._exit(code)
