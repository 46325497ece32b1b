"""World-size-2 gloo tests (CPU) of the data-parallel host logic in doodle_b200/dist.py: sharding the sun
batch and reducing the packed metric vector must reproduce the single-process means and gradients."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import helio_oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case():
    rng = np.random.default_rng(3)
    N, B, R = 6, 8, 16
    helio = np.concatenate([rng.random((N, 2)) * 10, np.zeros((N, 1))], 1).astype(np.float32)
    d = np.array([[0.5, 0.5, 0.7071]]) + 0.02 * rng.standard_normal((B, 3))
    sun = (d / np.linalg.norm(d, axis=1, keepdims=True) * 14142.0).astype(np.float32)
    ideal = orc.calculate_ideal_normals(sun, helio, [0., -5., 0.])
    act = (ideal + 0.02 * rng.standard_normal(ideal.shape)).astype(np.float32)
    errs = (rng.standard_normal((B, N, 2)) * 60).astype(np.float32)
    dm = (rng.random((B, R, R)) * 5).astype(np.float32)
    return helio, sun, act, errs, dm, R


def _local_sums(lo, hi):
    """Per-rank packed sums {sq, dist, bound, angle} and their gradients, from the oracle on the rank's slice."""
    helio, sun, act, errs, dm, R = _case()
    sl = slice(lo, hi)
    B, N = hi - lo, helio.shape[0]
    m, mon, _, _ = orc.env_step(sun[sl], act[sl], errs[sl], helio, [0., -5., 0.], [0., 1., 0.], (15., 15.), R, 0.1, dm[sl], weights=(0, 0, 0, 0))
    sums = np.array([m["mse"] * B * R * R, m["dist"] * B, m["bound"] * B * N, m["alignment_loss"] * B * N], np.float64)
    return sums, mon["mae_image"].reshape(-1)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from doodle_b200.dist import all_gather_cat, global_means, global_quantile, shard_bounds
    helio, sun, act, errs, dm, R = _case()
    B, N = sun.shape[0], helio.shape[0]
    lo, hi = shard_bounds(B, rank, world)
    sums, mae = _local_sums(lo, hi)
    x = torch.tensor(sums, dtype=torch.float64, requires_grad=True)
    inv = 1.0 / torch.tensor([B * R * R, B, B * N, B * N], dtype=torch.float64)
    means = global_means(x, inv)
    (means * torch.tensor([1., 2., 3., 4.], dtype=torch.float64)).sum().backward()
    cutoff = global_quantile(torch.tensor(mae), 0.8)
    gathered = all_gather_cat(torch.tensor(mae))
    q.put((rank, means.detach().numpy(), x.grad.numpy(), float(cutoff), gathered.numpy()))
    dist.destroy_process_group()


def test_sharded_metrics_equal_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    helio, sun, act, errs, dm, R = _case()
    B, N = sun.shape[0], helio.shape[0]
    m, mon, _, _ = orc.env_step(sun, act, errs, helio, [0., -5., 0.], [0., 1., 0.], (15., 15.), R, 0.1, dm, weights=(0, 0, 0, 0))
    ref = np.array([m["mse"], m["dist"], m["bound"], m["alignment_loss"]], np.float64)
    inv = 1.0 / np.array([B * R * R, B, B * N, B * N], np.float64)
    for rank, means, grad, cutoff, gathered in res:
        np.testing.assert_allclose(means, ref, rtol=1e-5)
        np.testing.assert_allclose(grad, np.array([1., 2., 3., 4.]) * inv, rtol=1e-12)   # d(global mean)/d(local sum)
        np.testing.assert_allclose(gathered, mon["mae_image"].reshape(-1), rtol=1e-5)
        np.testing.assert_allclose(cutoff, np.quantile(mon["mae_image"].reshape(-1).astype(np.float64), 0.8), rtol=1e-5)


def test_shard_bounds():
    from doodle_b200.dist import shard, shard_bounds
    assert [shard_bounds(8, r, 4) for r in range(4)] == [(0, 2), (2, 4), (4, 6), (6, 8)]
    with pytest.raises(ValueError):
        shard_bounds(10, 0, 4)
    t = torch.arange(12).view(6, 2)
    assert torch.equal(torch.cat([shard(t, r, 3) for r in range(3)]), t)


class _FakeEnv:
    """Stands in for HelioEnv on CPU: same hooks make_sharded_env overrides."""

    def __init__(self, batch_size, **kw):
        self.batch_size = batch_size
        self.error_mask_ratio = 0.2
        self._inv_counts = 1.0 / torch.tensor([float(batch_size)] * 4)
        self.sun_pos = self._sample_sun_pos()

    def _sample_sun_pos(self):
        return torch.rand(self.batch_size, 3)


def test_make_sharded_env_slices_global_sun_batch():
    from doodle_b200.dist import make_sharded_env
    torch.manual_seed(9)
    full = _FakeEnv(8).sun_pos
    parts = []
    for r in range(2):
        env = make_sharded_env(_FakeEnv, global_batch_size=8, rank=r, world_size=2, seed=9)
        assert env.batch_size == 4 and env.sun_pos.shape == (4, 3)
        parts.append(env.sun_pos)
    assert torch.equal(torch.cat(parts), full)


def test_sharded_randn_is_a_slice_of_the_global_draw():
    """Identically seeded ranks must hold disjoint slices of ONE global draw (not W copies of a local draw) and leave
    the generator where the single-process environment leaves it (dist.py seeded every rank identically and
    drew local-sized error tensors)."""
    from doodle_b200.dist import sharded_randn, shard_bounds
    G, N, W = 12, 5, 3
    torch.manual_seed(21)
    full = torch.randn(G, N, 2)
    after_full = torch.rand(4)
    parts = []
    for r in range(W):
        lo, hi = shard_bounds(G, r, W)
        torch.manual_seed(21)
        parts.append(sharded_randn((G, lo, hi), hi - lo, N, 2))
        assert torch.equal(torch.rand(4), after_full)            # generator state == single-process state
    assert torch.equal(torch.cat(parts), full)
    assert not torch.equal(parts[0], parts[1])                   # ranks differ
    torch.manual_seed(21)
    assert torch.equal(sharded_randn(None, G, N, 2), full)       # no shard: plain draw
    torch.manual_seed(21)
    assert sharded_randn((G, 0, 4), 1, N, 2).shape == (1, N, 2)  # not a batch-sized draw (B == 1 legacy errors): plain


def _minmax_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from doodle_b200.dist import all_reduce_minmax
    vals = torch.tensor([3.0, 7.0]) if rank == 0 else torch.tensor([-1.0, 5.0])
    mn, mx = all_reduce_minmax(vals.min(), vals.max())
    q.put((rank, float(mn), float(mx)))
    dist.destroy_process_group()


def test_ref_min_max_all_reduce():
    """ref_min / ref_max (test_environment.py:369-370) must be the extrema over the GLOBAL batch on every rank."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_minmax_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, mn, mx in res:
        assert (mn, mx) == (-1.0, 7.0)


class _FakeField:
    def __init__(self, batch_shard=None):
        self.batch_shard = batch_shard


class _FakeEnv2(_FakeEnv):
    """Also exercises the hooks HelioEnv reads: the class-level _batch_shard and _reduce_minmax."""

    def __init__(self, batch_size, **kw):
        self.field = _FakeField(batch_shard=getattr(self, "_batch_shard", None))
        super().__init__(batch_size, **kw)


def test_make_sharded_env_passes_the_shard_to_the_fields():
    from doodle_b200.dist import make_sharded_env
    env = make_sharded_env(_FakeEnv2, global_batch_size=8, rank=1, world_size=2, seed=9)
    assert env.field.batch_shard == (8, 4, 8)
    mn, mx = env._reduce_minmax(torch.tensor(1.0), torch.tensor(2.0))     # no process group: identity
    assert float(mn) == 1.0 and float(mx) == 2.0
