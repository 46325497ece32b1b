"""HelioEnv -- API mirror of the reference's Gym-style environment on top of the sm_100a kernels.

Mirrors ``test_environment.HelioEnv`` (reference test_environment.py:175-526): same constructor
signature and defaults, same RNG draw order at construction / reset, same ``reset`` / ``step``
return contracts (obs{'img','aux'}, metrics{'mse','dist','bound','alignment_loss'} with grad,
monitor{...}).  The setup-time pieces the scope table leaves in Python (sun-cone sampling, scipy
EDT distance maps) are restated here; the per-step work runs in K1-K4.

Differences, all opt-in or bug-compatible:
  * ``new_sun_pos_every_reset=True`` works (the reference calls a method that does not exist,
    test_environment.py:379);
  * the six NaN/Inf asserts (test_environment.py:495-501) read ONE fused triple {mse, dist, bound}.  ``check_finite=True``
    (default) copies it to pinned host memory without synchronising and raises the asserts of step t at the NEXT entry
    into the env (step / reset / set_sun_pos / ``flush_checks()``): a step never stalls the device, so the host can
    queue the backward while the forward still runs.  ``check_finite="sync"`` restores the reference's immediate
    asserts (one device sync per step); ``False`` switches them off;
  * ``fused_step`` (keyword-only, default True): ``step`` runs as one C-ABI call forward and one backward
    (helio_step_fwd / helio_step_bwd: the same kernels, far less host time for small fields); the
    error-mask and exponential-risk variants are formed from its per-image sums / per-heliostat bounds;
  * a HOST action (CPU tensor or np.ndarray, which the reference accepts too, :411-412) is copied in on a side
    stream while the target renders, and its gradient is returned in pinned host memory, copied out slice by slice
    under the backward kernels (functional.HostStepFn); ``obs['aux']`` and ``monitor['normals']`` are then built from
    the device copy, which HostStepFn returns as a differentiable output (gradients through them reach the host action);
  * ``cull`` (keyword-only, default False): the fused step contracts only over the heliostats whose footprint can
    reach the receiver (helio_cull: every dropped term is below 2^-40 of its peak on every pixel), which is most of the
    speed-up available when orientation errors are large; dense evaluation of every term, as in the reference, is the
    default;
  * ``graph`` (keyword-only, default "auto"): for small fields -- where a step is launch + Python overhead, not kernel
    time -- ``step`` replays captured CUDA graphs of the same fused forward / backward (graphs.StepGraph), transparently:
    metrics keep their autograd history, the caller's ``loss.backward()`` works, several steps may be alive at once.
    "auto" starts after two eager steps of a small shape (B*N*R^2 <= 2^33) on a device action without error mask,
    exponential risk, culling or sharding; True drops the size limit; False never replays.  Results are bit-identical.
    In graph mode the NaN/Inf asserts of step t are raised at the start of the next step / reset (no sync per step);
  * ``action_space="angular"`` (keyword-only): two angles per heliostat instead of a normal, the action space of the
    reference's experimental newenv/test_environment_angular.py (``angles_to_normals``);
  * encoder feed from the splat epilogue (keyword-only): ``HelioEnv(com=True)`` adds ``monitor['com']`` = the centre of
    mass of every image (layers/center_of_mass.py:21-60), differentiable; ``step(action, img_out=hist[:, -1])`` also writes
    the image straight into a caller's history slot (train_with_env.py:207-209).  Both come out of the kernel that renders
    the image, not from extra passes over it;
  * ``cache_target`` (keyword-only extension, default True).  The reference re-renders the target of the
    error-free field every step (test_environment.py:429-435) although it depends on ``sun_pos`` only; the
    cached image is bit-identical to the re-rendered one.  The cache is keyed on the ``sun_pos`` tensor's
    storage and version counter, so ``set_sun_pos`` AND in-place edits of ``env.sun_pos`` both invalidate it;
    ``cache_target=False`` restores the per-step target render.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F
from scipy.ndimage import distance_transform_edt

from .field import HelioField
from .functional import HostStepFn, ImageLossFn, StepFn, _cf, image_max, profiling, require_cuda
from .functional import distance_maps as _distance_maps_cuda
from .graphs import GraphStepFn, StepGraph

try:  # gymnasium is optional: only Env / spaces.Box / spaces.Dict are touched (test_environment.py:11-12)
    import gymnasium as gym
    from gymnasium import spaces
    _EnvBase = gym.Env
except Exception:  # pragma: no cover - gymnasium is not installed in the build image
    gym = None
    _EnvBase = object

    class _Box:
        def __init__(self, low=None, high=None, shape=None, dtype=None):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

    class _Dict(dict):
        def __init__(self, d):
            super().__init__(d)

    class spaces:  # noqa: N801 - mirrors the gymnasium module name
        Box = _Box
        Dict = _Dict


# ----------------------------------------------------------------------------------------------
# host-side helpers (setup time; device-agnostic so they are testable without a GPU)
# ----------------------------------------------------------------------------------------------
def azimuth_elevation_to_primary_direction(azimuth_deg: float, elevation_deg: float, device=None) -> torch.Tensor:
    """Unit direction for (azimuth, elevation) in degrees (test_environment.py:18-40)."""
    az = math.radians(azimuth_deg)
    el = math.radians(elevation_deg)
    vec = torch.tensor([math.cos(el) * math.cos(az), math.cos(el) * math.sin(az), math.sin(el)],
                       dtype=torch.float32, device=device)
    return vec / torch.norm(vec)


def sample_cone_directions(n: int, axis: torch.Tensor, half_angle_deg: float, device=None,
                           force_upper_hemisphere: bool = False) -> torch.Tensor:
    """n unit vectors uniform on the spherical cap around ``axis`` (test_environment.py:42-88).

    Draw order (two torch.rand(n) calls: cos-theta first, then phi) matches the reference so that a
    seeded run samples the same suns."""
    device = device or axis.device
    a = F.normalize(axis.to(device), dim=0)
    alpha = math.radians(half_angle_deg)
    helper = torch.tensor([0.0, 0.0, 1.0], device=device)
    if torch.abs(a[2]) > 0.999:
        helper = torch.tensor([0.0, 1.0, 0.0], device=device)
    u = F.normalize(torch.linalg.cross(helper, a), dim=0)
    v = torch.linalg.cross(a, u)
    u01 = torch.rand(n, device=device)
    cos_theta = 1.0 - u01 * (1.0 - math.cos(alpha))
    sin_theta = torch.sqrt(torch.clamp(1.0 - cos_theta ** 2, min=0.0))
    phi = 2.0 * math.pi * torch.rand(n, device=device)
    dirs = (u[None, :] * (sin_theta * torch.cos(phi))[:, None]
            + v[None, :] * (sin_theta * torch.sin(phi))[:, None]
            + a[None, :] * cos_theta[:, None])
    dirs = F.normalize(dirs, dim=1)
    if force_upper_hemisphere:
        dirs[:, 2] = torch.abs(dirs[:, 2])
    return dirs


def angles_to_normals(action: torch.Tensor, num_heliostats: int) -> torch.Tensor:
    """Angular action space of the reference's experimental env (newenv/test_environment_angular.py:205-214): two angles per
    heliostat -> mirror normals [B,N,3], by rotating a north-pointing normal (0,1,0) with ``rotate_normals_batch``
    (newenv_rl_test_multi_error.py:78-104): about Up by ``a[...,1]``, then about East by ``a[...,0]``.  As there, the
    angles go through the rotation helper's mrad scaling (x 1e-3 rad).  Plain differentiable torch ops (a few tiny
    kernels per step); the normals then enter the same fused step."""
    a = action.reshape(action.shape[0], num_heliostats, 2)
    e, u = a[..., 0] * 1e-3, a[..., 1] * 1e-3
    cu = torch.cos(u)
    return torch.stack([-torch.sin(u), torch.cos(e) * cu, torch.sin(e) * cu], dim=-1)


def make_distance_maps(imgs: torch.Tensor, thr: float = 0.5, impl: str = "auto") -> torch.Tensor:
    """Per-image Euclidean distance to the >thr*max region (test_environment.py:92-97).

    CUDA tensors go through the exact GPU transform (helio_distance_maps, bit-identical to scipy's exact EDT:
    no device->host->device round trip, no Python loop over B); ``impl="scipy"`` keeps the reference's host path
    (the only one available for CPU tensors, which the product never produces)."""
    if impl not in ("auto", "cuda", "scipy"):
        raise ValueError(f"unknown distance-map implementation {impl!r}")
    if impl == "cuda" or (impl == "auto" and imgs.is_cuda):
        return _distance_maps_cuda(imgs, thr)
    maps = []
    for img in imgs.detach().cpu().numpy():
        mask = (img > thr * img.max()).astype(np.uint8)
        maps.append(distance_transform_edt(1 - mask))
    return torch.tensor(np.stack(maps), dtype=torch.float32, device=imgs.device)


# ----------------------------------------------------------------------------------------------
class HelioEnv(_EnvBase):

    def __init__(self,
                 heliostat_pos,
                 targ_pos,
                 targ_area,
                 targ_norm,
                 sigma_scale=0.1,
                 error_scale_mrad=180.0,
                 initial_action_noise=0.0,
                 resolution=128,
                 batch_size=25,
                 device='cuda',
                 new_sun_pos_every_reset=False,
                 new_errors_every_reset=True,
                 use_error_mask=False,
                 error_mask_ratio=0.2,
                 exponential_risk=False,
                 single_sun=False,
                 azimuth=45.0,
                 elevation=45.0,
                 *,
                 cache_target=True,
                 check_finite=True,
                 fused_step=True,
                 distance_maps_impl="auto",
                 cull=False,
                 graph="auto",
                 action_space="normals",
                 com=False,
                 ):
        super().__init__()
        require_cuda(torch.device(device), "HelioEnv")

        if not isinstance(heliostat_pos, torch.Tensor):
            heliostat_pos = torch.tensor(heliostat_pos, dtype=torch.float32, device=device)
        if not isinstance(targ_pos, torch.Tensor):
            targ_pos = torch.tensor(targ_pos, dtype=torch.float32, device=device)
        if not isinstance(targ_norm, torch.Tensor):
            targ_norm = torch.tensor(targ_norm, dtype=torch.float32, device=device)

        self.resolution = resolution
        self.batch_size = batch_size
        self.device = device

        self.heliostat_pos = heliostat_pos
        self.num_heliostats = heliostat_pos.shape[0]
        self.targ_pos = targ_pos
        self.targ_area = targ_area
        self.targ_norm = targ_norm
        self.azimuth = azimuth
        self.elevation = elevation
        self.sigma_scale = sigma_scale
        self.error_scale_mrad = error_scale_mrad
        self.initial_action_noise = initial_action_noise
        self.sun_pos = None
        self.sun_errors = None
        self.new_sun_pos_every_reset = new_sun_pos_every_reset
        self.new_errors_every_reset = new_errors_every_reset
        self.single_sun = single_sun
        self.cache_target = cache_target
        self.check_finite = check_finite
        self.fused_step = fused_step
        self.cull = cull                               # footprint culling of the noisy render (helio_cull), off = dense
        self.host_chunks = "auto"                      # slices of the sun batch a host-resident action's step overlaps its copies in:
                                                       # an int, or "auto" = 8 for batches of >= 24 waves of tiles, else 4 (measured)
        self._copy_stream = None
        self.distance_maps_impl = distance_maps_impl   # "auto"/"cuda": GPU EDT; "scipy": the reference's host path
        self._target_cache = None
        if check_finite not in (True, False, "sync"):
            raise ValueError(f"check_finite must be True, False or 'sync' (got {check_finite!r})")
        self._finite_host = None                       # pinned {mse, dist, bound} of the last eager step (deferred asserts)
        self._finite_event = None
        self._finite_pending = False
        if graph not in (True, False, "auto"):
            raise ValueError(f"graph must be True, False or 'auto' (got {graph!r})")
        # encoder feed: com=True adds monitor['com'] [B,2] = CenterOfMass2D(obs['img']) (differentiable),
        # accumulated in the splat epilogue while the image tile is in registers (no extra pass over the image)
        self.com = bool(com)
        self.graph = graph                             # transparent CUDA-graph replay of step (graphs.StepGraph)
        self._step_graph = None
        self._graph_warm = 0

        if action_space not in ("normals", "angular"):
            raise ValueError(f"action_space must be 'normals' or 'angular' (got {action_space!r})")
        # "angular" (keyword-only extension): step takes two angles per heliostat, [B,2N] or [B,N,2], mapped to normals by
        # angles_to_normals (the action space of newenv/test_environment_angular.py:110-111,205-214); everything after
        # the mapping -- render, losses, monitors -- is the current env's step
        self.action_kind = action_space
        action_dim = heliostat_pos.shape[0] * (2 if action_space == "angular" else 3)
        self.action_space = spaces.Box(low=-1.0, high=1.0, shape=(action_dim,), dtype=np.float32)
        self.observation_space = spaces.Dict({
            'img': spaces.Box(low=0.0, high=np.inf, shape=(self.batch_size, resolution, resolution), dtype=np.float32),
            'aux': spaces.Box(low=-np.inf, high=np.inf, shape=(self.batch_size, 3 + self.heliostat_pos.shape[0] * 3),
                              dtype=np.float32),
        })

        # two fields, same order as the reference so that seeded error draws line up (:255-277)
        common = dict(heliostat_positions=self.heliostat_pos, target_position=self.targ_pos, target_area=self.targ_area,
                      target_normal=self.targ_norm, sigma_scale=self.sigma_scale, resolution=self.resolution,
                      max_batch_size=self.batch_size, device=self.device,
                      batch_shard=getattr(self, "_batch_shard", None))   # set by dist.make_sharded_env: global draws, sliced
        self.ref_field = HelioField(error_scale_mrad=0.0, **common)
        self.noisy_field = HelioField(error_scale_mrad=self.error_scale_mrad, **common)
        for f in (self.ref_field, self.noisy_field):
            f.set_boundary_geometry(self.targ_pos, self.targ_norm, self.targ_area)
        self.noisy_field.cull = cull                   # composed route (error-mask / exponential-risk variants, reset)

        self.use_error_mask = use_error_mask
        self.error_mask_ratio = error_mask_ratio
        self.exponential_risk = exponential_risk

        B, N, R = float(self.batch_size), float(self.num_heliostats), float(self.resolution)
        self._inv_counts = 1.0 / torch.tensor([B * R * R, B, B * N, B * N], dtype=torch.float32, device=self.device)

        self.set_sun_pos(self._sample_sun_pos())

    # ------------------------------------------------------------------ sun positions
    def _sample_sun_dirs(self) -> torch.Tensor:
        """Sun directions as sampled in the reference constructor (test_environment.py:286-321)."""
        half_angle_deg, force_upper = 2.0, True
        n = 1 if self.single_sun else self.batch_size
        if self.azimuth is not None and self.elevation is not None:
            primary = azimuth_elevation_to_primary_direction(self.azimuth, self.elevation, device=self.device)
            dirs = sample_cone_directions(n=n, axis=primary, half_angle_deg=half_angle_deg, device=self.device,
                                          force_upper_hemisphere=force_upper)
        else:
            dirs = F.normalize(torch.randn(n, 3, device=self.device), dim=1)
        if self.single_sun:
            dirs = dirs.repeat(self.batch_size, 1)
        if self.azimuth is None or self.elevation is None:
            dirs[:, 2] = torch.abs(dirs[:, 2])
        return dirs

    def _sample_sun_pos(self) -> torch.Tensor:
        radius = math.hypot(10000, 10000)                                   # :324
        return self._sample_sun_dirs() * radius

    def set_sun_pos_from_azimuth_elevation(self, azimuth_deg: float, elevation_deg: float, device=None):
        """Resample the suns around a new (azimuth, elevation) (the reference's version, :332-357, is
        unfinished: it never stores the result)."""
        self.azimuth, self.elevation = azimuth_deg, elevation_deg
        self.set_sun_pos(self._sample_sun_pos())

    def set_sun_pos(self, sun_positions: torch.Tensor):
        """Fix the current sun positions and rebuild target-dependent state (:359-370)."""
        self._check_deferred()
        self.sun_pos = sun_positions.clone().detach().to(device=self.device, dtype=torch.float32)
        self._target_cache = None
        self.ref_field.init_actions(self.sun_pos)
        with torch.no_grad():
            ideal_normals = self.ref_field.calculate_ideal_normals(self.sun_pos)
            timg, _ = self.ref_field.render(self.sun_pos, self.ref_field.initial_action, ideal_normals)
        self.distance_maps = make_distance_maps(timg, impl=self.distance_maps_impl)
        self.ref_min, self.ref_max = self._reduce_minmax(torch.min(timg), torch.max(timg))

    # ------------------------------------------------------------------ reset / step
    def reset(self):
        """Returns {'img': [B,R,R], 'aux': [B,3+3N] = cat(sun_pos, ideal normals)} (:372-400)."""
        self._check_deferred()
        if self.new_sun_pos_every_reset:
            self.set_sun_pos(self._sample_sun_pos())
        if self.new_errors_every_reset:
            self.noisy_field.reset_errors()
        with torch.no_grad():
            self.noisy_field.init_actions(self.sun_pos)
            out = self.noisy_field._render_full(self.sun_pos, self.noisy_field.initial_action, want_aux=True)
        self.ideal_normals = out.ideal
        aux = torch.cat([self.sun_pos, out.ideal.flatten(1)], dim=1)
        return {'img': out.img, 'aux': aux}

    # ---- exact target cache: (target, tx) depend on sun_pos only (:429-436) --------------------------------------
    def _cache_key(self):
        sp = self.sun_pos
        return (sp.data_ptr(), sp._version, self.ref_field.splat_impl)

    def _cached_target(self):
        """(target, tx) of the current suns or (None, None); an in-place edit of ``sun_pos`` bumps its version counter
        and a replaced tensor has another address, so a stale image is never returned."""
        c = self._target_cache
        if self.cache_target and c is not None and c[0] == self._cache_key():
            return c[1], c[2]
        return None, None

    def _store_target(self, target, tx):
        if self.cache_target:
            self._target_cache = (self._cache_key(), target, tx)

    def _target(self, ideal: torch.Tensor):
        """Target image of the error-free field and its per-image max (:429-436)."""
        target, tx = self._cached_target()
        if target is not None:
            return target, tx
        with torch.no_grad():
            target = self.ref_field._render_full(self.sun_pos, ideal, want_aux=False).img
            tx = image_max(target)
        self._store_target(target, tx)
        return target, tx

    def _quantile_cutoff(self, avg_error_per_heatmap: torch.Tensor) -> torch.Tensor:
        """Global quantile over the batch (:445); the sharded env overrides this with an all-gather."""
        return torch.quantile(avg_error_per_heatmap, 1 - self.error_mask_ratio)

    def _reduce_minmax(self, mn: torch.Tensor, mx: torch.Tensor):
        """ref_min / ref_max of the target images (:369-370); the sharded env all-reduces (min, max) here."""
        return mn, mx

    def _reduce_means(self, sums: torch.Tensor) -> torch.Tensor:
        """Packed {sum sq, sum dist, sum bound, sum angle} -> the four means (:128,455-457); the sharded
        env all-reduces here."""
        return sums * self._inv_counts

    # ---- transparent CUDA-graph replay (graphs.StepGraph) ---------------------------------------------------------
    def _check_deferred(self):
        """The NaN/Inf asserts of step t (:495-501) are raised here, at the next entry into the env."""
        sg = getattr(self, "_step_graph", None)
        if sg is not None:
            sg.check_pending()
        if getattr(self, "_finite_pending", False):
            self._finite_pending = False
            self._finite_event.synchronize()            # long done by the time the caller comes back
            mse, dist_l, bound = self._finite_host.numpy().tolist()
            if not (mse - mse == 0.0 and dist_l - dist_l == 0.0 and bound - bound == 0.0):
                assert not math.isnan(mse), "MSE is NaN"
                assert not math.isnan(dist_l), "Distance loss is NaN"
                assert not math.isnan(bound), "Boundary loss is NaN"
                assert not math.isinf(mse), "MSE is Inf"
                assert not math.isinf(dist_l), "Distance loss is Inf"
                assert not math.isinf(bound), "Boundary loss is Inf"

    def _host_chunks(self, B: int) -> int:
        if self.host_chunks != "auto":
            return int(self.host_chunks)
        from .functional import _wave_quantum
        return 8 if B >= 24 * _wave_quantum(self.noisy_field.device) else 4

    def close(self):
        """Raise pending asserts and release the captured CUDA graphs (and the static buffers they own).  A sharded env
        whose graphs captured the NCCL all-reduce must be closed before ``torch.distributed.destroy_process_group()``:
        destroying a communicator while graphs that captured its kernels are alive hangs."""
        try:
            self._check_deferred()
        finally:
            self._step_graph = None
            self._graph_warm = 0

    def flush_checks(self):
        """Raise any pending NaN/Inf assert now (waits for the last step's forward)."""
        self._check_deferred()

    def _defer_finite(self, means: torch.Tensor):
        if self._finite_host is None:
            self._finite_host = torch.zeros(3, dtype=torch.float32, device="cpu").pin_memory()
            self._finite_event = torch.cuda.Event()
        self._finite_host.copy_(means.detach()[:3], non_blocking=True)
        self._finite_event.record()
        self._finite_pending = True

    def _graph_eligible(self, action) -> bool:
        if self.graph is False or not self.fused_step or self.cull or self.use_error_mask or self.exponential_risk or self.com:
            return False
        if self.check_finite == "sync":                 # immediate asserts need the eager step's device sync
            return False
        if not (isinstance(action, torch.Tensor) and action.is_cuda):
            return False
        cls = type(self)
        if cls._reduce_means is not HelioEnv._reduce_means and not self._graph_reduce_ok():
            return False                                # a reduction hook the graph cannot reproduce
        if self.graph == "auto":
            B, N, R = self.batch_size, self.num_heliostats, self.resolution
            if B * N * R * R > (1 << 33) or B * R * R > (1 << 26):
                return False                            # kernel time dominates: replay (and its arena clone) buys nothing
            if self._graph_warm < 2:                    # two eager steps first: one-time setup, cached target
                self._graph_warm += 1
                return False
        # a caller capturing its own graph, or timing kernels with helio_profile_*, gets the eager kernels
        return not profiling() and not torch.cuda.is_current_stream_capturing()

    def _graph_reduce_ok(self) -> bool:
        """Sharded env (dist.make_sharded_env): its packed all-reduce can be captured into the forward graph when the
        process group runs on NCCL (a gloo group lives on the host and cannot)."""
        spec = getattr(self, "_graph_reduce", None)
        if spec is None:
            return False
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_backend(spec[0]) == "nccl"

    def _graph_step(self, action):
        nf = self.noisy_field
        sg = self._step_graph
        impl_bwd = nf.splat_impl if nf.splat_impl_bwd is None else nf.splat_impl_bwd
        if sg is None or sg.impl != nf.splat_impl or sg.impl_bwd != impl_bwd or sg.scene is not nf.scene():
            sg = self._step_graph = StepGraph(self)
            sg.capture_forward()
        B, N = self.batch_size, self.num_heliostats
        if action.dim() not in (2, 3) or action.shape[0] != B or action.numel() != 3 * B * N:
            action = action.reshape(B, N, 3)                     # raises on a wrong element count, like the reference's view
        img, means, aux, refl, bounds, angles, mae, ideal = GraphStepFn.apply(action, sg)
        mse, dist_l, bound, alignment_loss = means.unbind(0)
        return ({'img': img, 'aux': aux},
                {'mse': mse, 'dist': dist_l, 'bound': bound, 'alignment_loss': alignment_loss},
                {'normals': action.view(B, -1, 3), 'reflected_rays': refl, 'ideal_normals': ideal, 'all_bounds': bounds,
                 'mae_image': mae, 'alignment_errors': angles})

    def step(self, action, *, img_out=None):
        """obs, metrics, monitor = step(action)   (:402-516).  action: [B, 3N] or [B, N, 3].
        ``img_out`` (keyword-only extension): a float32 [B,R,R] view (any batch stride) that also receives the image,
        e.g. ``hist[:, -1]`` of the rollout's history buffer (train_with_env.py:207-209)."""
        self._check_deferred()
        if img_out is not None and not (self.fused_step and isinstance(action, torch.Tensor) and action.is_cuda):
            raise ValueError("img_out needs the fused step and a device action")
        if self.action_kind == "angular":
            if isinstance(action, np.ndarray):
                action = torch.from_numpy(np.ascontiguousarray(action, dtype=np.float32))
            action = angles_to_normals(action.to(self.device), self.num_heliostats)
        if img_out is None and self._graph_eligible(action):
            return self._graph_step(action)
        B, N, R = self.batch_size, self.num_heliostats, self.resolution
        com = None
        fused = self.fused_step
        if isinstance(action, np.ndarray):                                           # :411-412
            action = torch.from_numpy(np.ascontiguousarray(action, dtype=np.float32))
        if not fused and action.device.type == "cpu":
            action = action.to(self.device)
        action_dev = action

        if fused and action.device.type == "cpu":
            # host-resident action: copies overlapped with the target render / the backward slices (HostStepFn)
            nf = self.noisy_field
            act = action if action.dtype == torch.float32 else action.float()
            cached = self._cached_target()
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=nf.device)
            img, packed, _actual, refl, ideal_normals, bounds, angles, per_img, target, tx, action_dev = HostStepFn.apply(
                act.contiguous(), self.sun_pos, _cf(nf._select_errors(B)), nf.heliostat_positions, self.distance_maps, nf.scene(),
                nf._geom_workspace(B), R, nf.splat_impl, nf.splat_impl if nf.splat_impl_bwd is None else nf.splat_impl_bwd,
                cached[0], cached[1], self._copy_stream, self._host_chunks(B), self.cull)
            if cached[0] is None:
                self._store_target(target, tx)
            out = SimpleNamespace(refl=refl, bounds=bounds, angles=angles)
        elif fused:
            nf = self.noisy_field
            act = action.to(device=nf.device, dtype=torch.float32)
            normals = act.reshape(B, N, 3).contiguous()
            cached = self._cached_target()
            img, packed, _actual, refl, ideal_normals, bounds, angles, per_img, target, tx, com, _com_sums = StepFn.apply(
                normals, self.sun_pos, _cf(nf._select_errors(B)), nf.heliostat_positions, self.distance_maps, nf.scene(),
                nf._geom_workspace(B), R, nf.splat_impl, nf.splat_impl if nf.splat_impl_bwd is None else nf.splat_impl_bwd,
                cached[0], cached[1], self.cull, self.com, img_out)
            if cached[0] is None:
                self._store_target(target, tx)
            out = SimpleNamespace(refl=refl, bounds=bounds, angles=angles)
        else:
            out = self.noisy_field._render_full(self.sun_pos, action, want_aux=True)     # K1 + K2
            img, ideal_normals = out.img, out.ideal
            target, tx = self._target(ideal_normals)
            per_img = ImageLossFn.apply(img, target, self.distance_maps, tx)             # K4: [B,3]
            packed = None
        aux = torch.cat([self.sun_pos.detach(), action_dev.flatten(1)], dim=1)
        avg_error_per_heatmap = per_img[:, 2] / float(R * R)

        if packed is None or self.use_error_mask or self.exponential_risk:
            # variants of the loss block (test_environment.py:445-452, :472-480) formed from the per-image sums and the
            # per-heliostat bounds; autograd carries their gradients back into the same backward kernels
            if self.use_error_mask:
                cutoff = self._quantile_cutoff(avg_error_per_heatmap)
                mask = (avg_error_per_heatmap > cutoff).float()
                sq, ds = per_img[:, 0] * mask, per_img[:, 1] * mask
            else:
                sq, ds = per_img[:, 0], per_img[:, 1]
            sums = packed[2:] if packed is not None else out.sums
            bound_sum = torch.exp(out.bounds + 1e-6).sum() if self.exponential_risk else sums[0]
            packed = torch.stack([sq.sum(), ds.sum(), bound_sum, sums[1]])
        means = self._reduce_means(packed)
        mse, dist_l, bound, alignment_loss = means.unbind(0)

        if self.check_finite is True:                                                # :495-501, raised at the next entry: no sync
            self._defer_finite(means)
        elif self.check_finite == "sync":                                            # the reference's immediate asserts, one sync
            if not bool(torch.isfinite(means[:3]).all()):
                assert not torch.isnan(mse).any(), "MSE is NaN"
                assert not torch.isnan(dist_l).any(), "Distance loss is NaN"
                assert not torch.isnan(bound).any(), "Boundary loss is NaN"
                assert not torch.isinf(mse).any(), "MSE is Inf"
                assert not torch.isinf(dist_l).any(), "Distance loss is Inf"
                assert not torch.isinf(bound).any(), "Boundary loss is Inf"

        metrics = {'mse': mse, 'dist': dist_l, 'bound': bound, 'alignment_loss': alignment_loss}
        obs = {'img': img, 'aux': aux}
        monitor = {
            'normals': action_dev.view(self.batch_size, -1, 3),      # host action: its device copy (autograd-linked)
            'reflected_rays': out.refl.view([-1, 3]),
            'ideal_normals': ideal_normals.view([-1, 3]),
            'all_bounds': out.bounds,
            'mae_image': avg_error_per_heatmap.view([-1, 1]),
            'alignment_errors': out.angles.detach().view([-1]),
        }
        if self.com:
            if com is None:                     # host-action / composed routes: the layer's own kernel (one pass over the image)
                from .layers import CenterOfMass2D
                com = CenterOfMass2D()(img)
            monitor['com'] = com
        return obs, metrics, monitor

    def seed(self, seed=None):
        """torch + numpy seeds (:518-526)."""
        if seed is not None:
            torch.manual_seed(seed)
            np.random.seed(seed)
