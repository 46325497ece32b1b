"""HelioField -- API mirror of the reference's optics core on top of the sm_100a kernels.

Mirrors ``newenv_rl_test_multi_error.HelioField`` (reference newenv_rl_test_multi_error.py:154-415):
same constructor, attributes, RNG draw order, return arity and shapes.  ``render`` runs K1
(helio_geom_fwd) + K2 (helio_splat_fwd) through autograd Functions whose backward is K3 + K1'
(helio_splat_bwd, helio_geom_bwd); no [B,N,R,R] tensor is ever materialised.  CUDA only.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

from . import _lib
from .dist import sharded_randn
from .functional import GeomFn, SplatFn, SPLAT_AUTO, Scene, _cf, _stream, require_cuda


class HelioField:
    """Heliostat field with per-sun-position error sampling (reference :154-160)."""

    def __init__(
        self,
        heliostat_positions: torch.Tensor,
        target_position: torch.Tensor,
        target_area: tuple,
        target_normal: torch.Tensor,
        error_scale_mrad: float = 1.0,
        sigma_scale: float = 0.01,
        initial_action_noise: float = 0.01,
        resolution: int = 100,
        device: torch.device | str = "cpu",
        max_batch_size: int = 25,
        *,
        batch_shard=None,
    ) -> None:
        # batch_shard (keyword-only extension, set by dist.make_sharded_env): (global_batch, lo, hi).  Random tensors with
        # a leading batch dimension are then drawn for the GLOBAL batch and this rank's rows kept (dist.sharded_randn).
        self.batch_shard = batch_shard
        self.device = torch.device(device)
        require_cuda(self.device, "HelioField")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        _lib.load()  # fail loudly now if the extension is missing
        self.max_batch_size = int(max_batch_size)

        f32 = dict(dtype=torch.float32, device=self.device)
        self.heliostat_positions = torch.as_tensor(heliostat_positions, **f32).contiguous()
        self.num_heliostats = self.heliostat_positions.shape[0]
        self.target_position = torch.as_tensor(target_position, **f32)
        self.target_width, self.target_height = target_area
        self.target_normal = torch.as_tensor(target_normal, **f32)
        self.target_normal = self.target_normal / self.target_normal.norm().clamp_min(1e-9)   # :192

        self.error_scale_mrad = float(error_scale_mrad)
        self.initial_action_noise = float(initial_action_noise)
        self.sigma_scale = float(sigma_scale)
        self.resolution = int(resolution)

        self.reset_errors()                                                                    # :203

        # basis on the target plane (:206-213)
        self.plane_u = torch.tensor([1.0, 0.0, 0.0], device=self.device)
        if torch.allclose(self.target_normal, torch.tensor([0.0, 1.0, 0.0], device=self.device)):
            self.plane_v = torch.tensor([0.0, 0.0, 1.0], device=self.device)
        else:
            v = torch.linalg.cross(self.target_normal, self.plane_u)
            self.plane_v = v / v.norm().clamp_min(1e-9)

        self.initial_action = None
        self.splat_impl = SPLAT_AUTO        # forward splat: SPLAT_AUTO | SPLAT_SIMT | SPLAT_TC
        self.splat_impl_bwd = None          # backward splat; None = same selector as forward
        self.cull = False                   # opt-in footprint culling in render() (cull.cuh); dense by default
        self._scene = None
        self._bnd = None
        self._workspace = {}

    # ------------------------------------------------------------------ C-ABI plumbing
    def set_boundary_geometry(self, targ_pos, targ_norm, targ_area, east=(1.0, 0.0, 0.0), up=(0.0, 0.0, 1.0)):
        """Constants HelioEnv.step passes to boundary() (test_environment.py:460-470)."""
        self._bnd = (tuple(float(x) for x in torch.as_tensor(targ_pos).flatten().tolist()),
                     tuple(float(x) for x in torch.as_tensor(targ_norm).flatten().tolist()),
                     (float(targ_area[0]), float(targ_area[1])), tuple(east), tuple(up))
        self._scene = None

    def scene(self) -> Scene:
        if self._scene is None:
            s = Scene()
            s.target_pos[:] = self.target_position.tolist()
            s.target_normal[:] = self.target_normal.tolist()
            s.plane_u[:] = self.plane_u.tolist()
            s.plane_v[:] = self.plane_v.tolist()
            s.width, s.height = float(self.target_width), float(self.target_height)
            s.sigma_scale = self.sigma_scale
            bnd = self._bnd or (tuple(self.target_position.tolist()), tuple(self.target_normal.tolist()),
                                (float(self.target_width), float(self.target_height)), (1.0, 0.0, 0.0), (0.0, 0.0, 1.0))
            s.bnd_targ_pos[:] = bnd[0]
            s.bnd_targ_norm[:] = bnd[1]
            s.bnd_width, s.bnd_height = bnd[2]
            s.bnd_u[:] = bnd[3]
            s.bnd_v[:] = bnd[4]
            self._scene = s
        return self._scene

    def _geom_workspace(self, B: int) -> torch.Tensor:
        """Ticket counter + block partials of K1's ordered reduction.  One workspace per (B, stream): two steps of the
        same field enqueued on different streams must not share the counter (include/helio_b200.h: one stream per
        workspace at a time)."""
        key = (B, _stream())
        ws = self._workspace.get(key)
        if ws is None:
            nbytes = _lib.load().helio_geom_workspace_bytes(B, self.num_heliostats)
            ws = torch.zeros((nbytes + 3) // 4, dtype=torch.int32, device=self.device)
            self._workspace[key] = ws
        return ws

    # --------------------------------------------------------------------- API (reference :220-304)
    def reset_errors(self) -> None:
        """Regenerate the [N,2] legacy and [max_B,N,2] batched error tensors (same draw order, :231-239)."""
        self.error_angles_mrad = (
            torch.randn(self.num_heliostats, 2, device=self.device) * self.error_scale_mrad
        )
        if self.max_batch_size >= 1:
            self.batch_error_angles_mrad = self._sample_error_angles(self.max_batch_size)
        else:
            self.batch_error_angles_mrad = None

    def _sample_error_angles(self, batch_size: int) -> torch.Tensor:
        return (
            sharded_randn(self.batch_shard, batch_size, self.num_heliostats, 2, device=self.device)
            * self.error_scale_mrad
        )

    def calculate_ideal_normals(self, sun_position: torch.Tensor) -> torch.Tensor:
        """Per-heliostat normals that hit the target (:256-278).  Setup-time helper; the hot path gets
        the same vectors fused out of K1."""
        sun = torch.as_tensor(sun_position, dtype=torch.float32, device=self.device)
        single = sun.dim() == 1
        s = sun.view(-1, 1, 3)
        helios = self.heliostat_positions.view(1, self.num_heliostats, 3)
        incidents = s - helios
        reflected = self.target_position.view(1, 1, 3) - helios
        inc_dir = incidents / incidents.norm(dim=2, keepdim=True).clamp_min(1e-9)
        ref_dir = reflected / reflected.norm(dim=2, keepdim=True).clamp_min(1e-9)
        normals = inc_dir + ref_dir
        normals = normals / normals.norm(dim=2, keepdim=True).clamp_min(1e-9)
        return normals[0] if single else normals

    def init_actions(self, sun_position: torch.Tensor) -> None:
        """ideal + N(0, initial_action_noise), renormalised (:291-304)."""
        ideal = self.calculate_ideal_normals(sun_position)
        if ideal.dim() == 3:
            noise = sharded_randn(self.batch_shard, ideal.shape[0], self.num_heliostats, 3, device=self.device) * self.initial_action_noise
        else:
            noise = torch.randn_like(ideal) * self.initial_action_noise
        noisy = ideal + noise
        if ideal.dim() == 2:
            noisy = noisy / noisy.norm(dim=1, keepdim=True).clamp_min(1e-9)
            self.initial_action = noisy.flatten()
        else:
            normed = noisy / noisy.norm(dim=2, keepdim=True).clamp_min(1e-9)
            self.initial_action = normed.view(ideal.shape[0], -1)

    # ------------------------------------------------------------------ render
    def _select_errors(self, B: int) -> torch.Tensor:
        """Error tensor render() uses for a batch of B suns (:340-353)."""
        if B == 1:
            return self.error_angles_mrad.unsqueeze(0)
        if self.batch_error_angles_mrad is not None and B <= self.batch_error_angles_mrad.shape[0]:
            return self.batch_error_angles_mrad[:B]
        return self._sample_error_angles(B)     # fresh draw every call, as in the reference

    def _render_full(self, sun: torch.Tensor, action: torch.Tensor, want_aux: bool, errs=None) -> SimpleNamespace:
        """K1 + K2 for sun [B,3]; returns img, actual, refl and (want_aux) ideal, bounds, angles, sums."""
        B = sun.shape[0]
        N = self.num_heliostats
        act = action if isinstance(action, torch.Tensor) else torch.as_tensor(action)
        act = act.to(device=self.device, dtype=torch.float32)
        if act.dim() == 1:
            act = act.unsqueeze(0)
        normals = act.reshape(B, N, 3).contiguous()
        if errs is None:
            errs = self._select_errors(B)
        errs = _cf(errs)
        ws = self._geom_workspace(B) if want_aux else None
        params, actual, refl, ideal, bounds, angles, sums = GeomFn.apply(
            normals, _cf(sun), errs, self.heliostat_positions, self.scene(), ws, want_aux)
        img = SplatFn.apply(params, self.resolution, float(self.target_width), float(self.target_height), self.splat_impl,
                            self.splat_impl if self.splat_impl_bwd is None else self.splat_impl_bwd, self.cull)
        return SimpleNamespace(img=img, actual=actual, refl=refl, ideal=ideal, bounds=bounds, angles=angles, sums=sums,
                               normals=normals)

    def render(
        self,
        sun_position: torch.Tensor,
        action: torch.Tensor,
        ideal_normals: torch.Tensor,
        show_spillage: bool = False,   # kept for API completeness (unused in the reference too)
        monitor: bool = False,
    ):
        """Irradiance image(s) on the target plane (:308-415).

        sun [3] -> ([R,R], actual [1,N,3][, refl [N,3]]);  sun [B,3] -> ([B,R,R], [B,N,3][, [B*N,3]]).
        ``ideal_normals`` is accepted and ignored, exactly like the reference (its only use is commented
        out at :365).
        """
        sun = torch.as_tensor(sun_position, dtype=torch.float32, device=self.device)
        batched = sun.dim() > 1
        if not batched:
            sun = sun.unsqueeze(0)
        out = self._render_full(sun, action, want_aux=False)
        images = out.img
        if not monitor:
            return (images[0], out.actual) if not batched else (images, out.actual)
        return (images[0], out.actual, out.refl) if not batched else (images, out.actual, out.refl)
