#!/usr/bin/env python3
"""BASELINE.json configs[2]: a policy trained through HelioEnv.step on the sm_100a renderer (N=50, 128x128, B=256).

The reference's trainer (train_with_env.py: CNN encoder + LSTM policy, rollout of T steps over a k-frame history,
alignment-loss pretraining, AdamP) is a CALLER of the hot path and is not rebuilt here; it cannot travel to the GPU box
either (it needs adamp / mlflow / plotly, none installed).  This is a compact stand-in with the same data flow
(rollout, train_with_env.py:171-216: reset -> [policy(hist, aux) -> env.step(normals)] x T -> loss.backward()), used to
show the environment training a network end to end and to measure env-steps/s with a policy in the loop.

    python examples/train_policy_c3.py [--iters 60] [--B 256] [--N 50] [--R 128] [--T 4] [--k 4] [--json out.json]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
import torch.nn.functional as F
from doodle_b200 import HelioEnv


class Policy(nn.Module):
    """(B,k,1,H,W) history + (B,aux) -> unit normals (B,N,3); conv encoder per frame, LSTM over frames, MLP head."""

    def __init__(self, num_heliostats, aux_dim, enc_dim=128, hid=128):
        super().__init__()
        self.num_h = num_heliostats
        self.enc = nn.Sequential(nn.Conv2d(1, 16, 5, stride=2, padding=2), nn.GELU(), nn.Conv2d(16, 32, 3, stride=2, padding=1), nn.GELU(),
                                 nn.Conv2d(32, 64, 3, stride=2, padding=1), nn.GELU(), nn.AdaptiveAvgPool2d(1), nn.Flatten(),
                                 nn.Linear(64, enc_dim))
        self.rnn = nn.LSTM(enc_dim, hid, batch_first=True)
        self.head = nn.Sequential(nn.LayerNorm(hid + aux_dim), nn.Linear(hid + aux_dim, 256), nn.GELU(), nn.Linear(256, num_heliostats * 3))

    def forward(self, img_seq, aux, hx=None):
        B, T = img_seq.shape[:2]
        e = self.enc(img_seq.flatten(0, 1)).view(B, T, -1)
        out, hx = self.rnn(e, hx)
        n = self.head(torch.cat([out[:, -1], aux], dim=1)).view(B, self.num_h, 3)
        return F.normalize(n, dim=2), hx


def rollout(env, policy, k, T):
    with torch.no_grad():
        s = env.reset()
    img, aux = s["img"], s["aux"]
    B, R = env.batch_size, env.resolution
    hist = torch.zeros(B, k, R, R, device=img.device)
    hist[:, -1] = img
    hx, losses = None, None
    for _ in range(T):
        normals, hx = policy(hist.unsqueeze(2).detach(), aux.detach(), hx)
        s, losses, _ = env.step(normals)
        hist = torch.roll(hist, -1, dims=1)
        hist[:, -1] = s["img"].detach()
    return losses


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=60); ap.add_argument("--B", type=int, default=256)
    ap.add_argument("--N", type=int, default=50); ap.add_argument("--R", type=int, default=128)
    ap.add_argument("--T", type=int, default=4); ap.add_argument("--k", type=int, default=4)
    ap.add_argument("--lr", type=float, default=2e-4); ap.add_argument("--json", default=None)
    a = ap.parse_args()
    dev = "cuda:0"
    torch.manual_seed(0)
    helio = torch.rand(a.N, 3, device=dev) * 10 + 80                              # train_with_env.py:227-230
    helio[:, 2] = 0
    env = HelioEnv(helio, torch.tensor([0., -5., 0.], device=dev), (15., 15.), torch.tensor([0., 1., 0.], device=dev), sigma_scale=0.01,
                   error_scale_mrad=90.0, resolution=a.R, batch_size=a.B, device=dev, new_errors_every_reset=True)
    policy = Policy(a.N, 3 + 3 * a.N).to(dev)
    opt = torch.optim.AdamW(policy.parameters(), lr=a.lr)
    hist = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for it in range(a.iters):
        losses = rollout(env, policy, a.k, a.T)
        loss = losses["alignment_loss"]                                           # the reference's current schedule (train_with_env.py:347-355)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(policy.parameters(), 1.0)
        opt.step()
        hist.append({k: float(v) for k, v in losses.items()})
        if it == 4:                                                               # steady-state clock starts after warm-up
            torch.cuda.synchronize(); t1 = time.perf_counter()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t1
    steps = (a.iters - 5) * a.T
    out = dict(config=f"N={a.N} R={a.R} B={a.B} T={a.T} k={a.k} (BASELINE.json configs[2] shape, stand-in policy)",
               env_steps_per_s=steps / dt, rollouts_per_s=(a.iters - 5) / dt, ms_per_env_step_incl_policy=dt / steps * 1e3,
               alignment_loss_first=hist[0]["alignment_loss"], alignment_loss_last=hist[-1]["alignment_loss"],
               alignment_loss_min=min(h["alignment_loss"] for h in hist), mse_last=hist[-1]["mse"], wall_s=time.perf_counter() - t0)
    print(json.dumps(out))
    if a.json:
        json.dump(dict(summary=out, history=hist), open(a.json, "w"), indent=1)
    return out


if __name__ == "__main__":
    main()
