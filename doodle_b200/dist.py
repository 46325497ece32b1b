"""Data-parallel sharding of the sun/episode batch across GPUs.

Every sun position b is independent in render forward/backward, so the global batch is split into
contiguous B/W slices, one process per GPU, heliostat geometry replicated.  The only exchange is
one all-reduce of the packed 4-float metric vector {sum sq, sum dist, sum bound, sum angle} per
step (NCCL over NVLink on GPUs, gloo in the CPU tests) plus, with ``use_error_mask``, an all-gather
of the B/W per-image errors for the global quantile (test_environment.py:445).  Images, monitors
and action gradients stay rank-local.

This module is host logic only (it never touches the kernels), so it is covered by world_size-2
gloo tests on CPU.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(global_batch: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[lo, hi) of the contiguous slice of suns owned by ``rank``; the batch must divide evenly so
    that the mean of local means equals the global mean."""
    if global_batch % world_size != 0:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world_size}")
    per = global_batch // world_size
    return rank * per, (rank + 1) * per


def shard(t: torch.Tensor, rank: int, world_size: int) -> torch.Tensor:
    lo, hi = shard_bounds(t.shape[0], rank, world_size)
    return t[lo:hi]


class _AllReduceSum(torch.autograd.Function):
    """y = sum over ranks of x.  Each rank's loss is the same global scalar L(y); autograd on rank r
    only needs dL/dx_r = dL/dy, so backward is the identity (no second collective)."""

    @staticmethod
    def forward(ctx, x, group):
        y = x.detach().clone()
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
        return y

    @staticmethod
    def backward(ctx, g):
        return g, None


def all_reduce_sum(x: torch.Tensor, group=None) -> torch.Tensor:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return x
    return _AllReduceSum.apply(x, group)


def all_gather_cat(x: torch.Tensor, group=None) -> torch.Tensor:
    """Concatenate a 1-D tensor over ranks (no gradient: used for the quantile cutoff only)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return x.detach()
    parts = [torch.empty_like(x) for _ in range(dist.get_world_size(group))]
    dist.all_gather(parts, x.detach().contiguous(), group=group)
    return torch.cat(parts)


def global_means(local_sums: torch.Tensor, inv_global_counts: torch.Tensor, group=None) -> torch.Tensor:
    """Packed local sums -> global means with ONE all-reduce; differentiable w.r.t. local_sums."""
    return all_reduce_sum(local_sums, group) * inv_global_counts


def global_quantile(local_values: torch.Tensor, q: float, group=None) -> torch.Tensor:
    return torch.quantile(all_gather_cat(local_values, group), q)


def sharded_randn(shard: Optional[Tuple[int, int, int]], batch: int, *tail: int, device=None) -> torch.Tensor:
    """``torch.randn(batch, *tail)`` for a batch-leading tensor of a sharded environment.

    ``shard = (global_batch, lo, hi)``: when ``batch`` is this rank's slice size the GLOBAL tensor
    ``[global_batch, *tail]`` is drawn and rows ``[lo:hi]`` are kept, so that identically seeded ranks (a) hold disjoint
    slices of one global draw -- not W copies of the same local draw -- and (b) leave the generator in the state the
    single-process environment leaves it in (a CUDA ``randn`` cannot be sliced by skipping ahead: which Philox counter
    feeds which element depends on the launch grid, i.e. on the total element count).  ``shard=None`` is a plain draw."""
    if shard is not None and batch == shard[2] - shard[1]:
        return torch.randn(shard[0], *tail, device=device)[shard[1]:shard[2]].contiguous()
    return torch.randn(batch, *tail, device=device)


def all_reduce_minmax(mn: torch.Tensor, mx: torch.Tensor, group=None):
    """Global (min, max) of rank-local extrema (HelioEnv.ref_min / ref_max, test_environment.py:369-370)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return mn, mx
    mn, mx = mn.detach().clone(), mx.detach().clone()
    dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    return mn, mx


def make_sharded_env(env_cls, *args, global_batch_size: int, rank: Optional[int] = None,
                     world_size: Optional[int] = None, group=None, seed: Optional[int] = None, **kwargs):
    """Build the rank-local slice of a ``global_batch_size`` HelioEnv.

    All ranks seed identically (``seed=`` or the caller's ``torch.manual_seed``) and every random tensor with a leading
    batch dimension -- sun directions, ``batch_error_angles_mrad`` (constructor and every ``reset_errors``), the
    ``init_actions`` noise -- is drawn at the GLOBAL batch size and sliced ``[lo:hi]`` (``sharded_randn``), so that
    concatenating the ranks reproduces the single-process environment exactly, draw for draw.  ``ref_min`` / ``ref_max``
    are all-reduced (min / max) in ``set_sun_pos``.  Metrics returned by ``step`` are the global means (one packed
    all-reduce); gradients w.r.t. the local actions are those of the global means.

    Small fields replay CUDA graphs that contain the NCCL all-reduce (``graph="auto"``): call ``env.close()`` before
    ``torch.distributed.destroy_process_group()``.
    """
    rank = dist.get_rank(group) if rank is None else rank
    world_size = dist.get_world_size(group) if world_size is None else world_size
    lo, hi = shard_bounds(global_batch_size, rank, world_size)

    class ShardedEnv(env_cls):  # type: ignore[misc, valid-type]
        _batch_shard = (global_batch_size, lo, hi)       # read by HelioEnv.__init__ -> HelioField(batch_shard=...)

        def _sample_sun_pos(self):
            # draw the global batch with the local batch size temporarily widened
            local = self.batch_size
            self.batch_size = global_batch_size
            try:
                full = super()._sample_sun_pos()
            finally:
                self.batch_size = local
            return full[lo:hi]

        def _reduce_means(self, sums):
            return global_means(sums, self._inv_counts / world_size, group)

        def _quantile_cutoff(self, avg):
            return global_quantile(avg, 1 - self.error_mask_ratio, group)

        def _reduce_minmax(self, mn, mx):
            return all_reduce_minmax(mn, mx, group)

        # transparent graph replay (graphs.StepGraph): the packed all-reduce is captured INTO the forward graph (NCCL
        # collectives are capturable), so a sharded small field replays one graph per step like the single-process env
        _graph_reduce = (group, world_size)

    if seed is not None:
        torch.manual_seed(seed)
    env = ShardedEnv(*args, batch_size=hi - lo, **kwargs)
    env.global_batch_size, env.rank, env.world_size = global_batch_size, rank, world_size
    return env
