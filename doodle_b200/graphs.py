"""CUDA-graph capture of ``HelioEnv.step`` + backward.

The reference's test-time-compute loops call ``env.step(candidate); loss.backward(); opt.step()``
``fine_steps_per_t`` times per rollout step with fixed shapes
(train_with_env_com_trunc_advantage_ttt.py:291-312, fine_adjustment_sanity_check.py:133-141).  For a small
field that loop is pure launch + Python overhead.  Every entry point of libhelio_sm100.so only enqueues
kernels on the caller's stream and never allocates, so the whole step (8 forward + 3 backward kernels plus
the handful of torch ops around them) can be captured once and replayed with one graph launch.

    gstep = GraphedStep(env, objective=lambda m: m["dist"])
    for _ in range(fine_steps):
        loss, grad = gstep(candidate)        # grad = d objective / d candidate, static buffers
        ...

Replays read ``env.sun_pos``, ``env.distance_maps`` and the noisy field's error tensors through the
addresses captured, so in-place updates of those tensors are seen; *replacing* them (``set_sun_pos``,
``reset`` with ``new_errors_every_reset``) needs ``recapture()``.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch


def _default_objective(m: Dict[str, torch.Tensor]) -> torch.Tensor:
    return m["mse"] + m["dist"] + m["bound"] + m["alignment_loss"]


class GraphedStep:
    """One CUDA graph holding ``env.step(action)`` and the backward of ``objective(metrics)`` to ``action``."""

    def __init__(self, env, objective: Optional[Callable[[Dict[str, torch.Tensor]], torch.Tensor]] = None,
                 warmup: int = 3):
        self.env = env
        self.objective = objective or _default_objective
        B, N = env.batch_size, env.num_heliostats
        self.action = torch.zeros(B, N, 3, device=env.device, requires_grad=True)
        self._warmup = warmup
        self.graph = None
        self.recapture()

    def _run(self):
        obs, metrics, monitor = self.env.step(self.action)
        loss = self.objective(metrics)
        grad, = torch.autograd.grad(loss, self.action)
        return obs, metrics, monitor, loss, grad

    def recapture(self, action: Optional[torch.Tensor] = None):
        env = self.env
        if action is None:
            action = env.ideal_normals if getattr(env, "ideal_normals", None) is not None else None
        with torch.no_grad():
            if action is not None:
                self.action.copy_(action.detach().reshape(self.action.shape))
        check = env.check_finite
        env.check_finite = False                       # a device->host sync cannot be captured
        try:
            side = torch.cuda.Stream(device=self.action.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):              # warm-up off the default stream: one-time setup
                for _ in range(self._warmup):          # (workspaces, func attributes, cached scene) happens here
                    self._run()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            from . import functional as Fn
            l0 = Fn.launch_count()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.obs, self.metrics, self.monitor, self.loss, self.grad = self._run()
            self.helio_kernels_per_replay = Fn.launch_count() - l0
        finally:
            env.check_finite = check

    def __call__(self, action: torch.Tensor):
        """Replays the step on ``action`` ([B,N,3] or [B,3N]); returns (loss, grad) living in static buffers that
        the next call overwrites (``self.obs`` / ``self.metrics`` / ``self.monitor`` likewise)."""
        with torch.no_grad():
            self.action.copy_(action.detach().reshape(self.action.shape))
        self.graph.replay()
        return self.loss, self.grad
