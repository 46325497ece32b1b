#!/usr/bin/env python3
"""Generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (build container only).

The reference (l3th4l/DOODLE) ships no golden vectors, so parity is pinned on outputs
of its own Python, imported read-only from /root/reference.  That tree does not exist
on the GPU box; the fixtures written here are committed and are what the tests read.

    python oracle/make_golden.py            # rewrites tests/golden/

``gymnasium`` is not installed in this image; test_environment.py only touches
``gym.Env``, ``spaces.Box`` and ``spaces.Dict`` (test_environment.py:11-12,175,241-252), so
an in-memory stub of those three names is injected before import.  Nothing from the
reference is copied: it is imported, called, and its outputs are saved.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("HELIO_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def import_reference():
    gym = types.ModuleType("gymnasium")
    spaces = types.ModuleType("gymnasium.spaces")

    class Env:  # noqa: D401 - stub
        pass

    class Box:
        def __init__(self, low=None, high=None, shape=None, dtype=None):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

    class Dict(dict):
        def __init__(self, d):
            super().__init__(d)

    gym.Env, gym.spaces = Env, spaces
    spaces.Box, spaces.Dict = Box, Dict
    sys.modules.setdefault("gymnasium", gym)
    sys.modules.setdefault("gymnasium.spaces", spaces)
    sys.path.insert(0, REF)
    import newenv_rl_test_multi_error as ref_field  # noqa: E402
    import test_environment as ref_env  # noqa: E402
    return ref_field, ref_env


def npy(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def render_case(ref_field, name, seed, N, R, B, helio_fn, sigma_scale, err_mrad, target_pos=(0., -5., 0.),
                target_normal=(0., 1., 0.), area=(15., 15.), single=False, tweak=None, sun_fn=None, w_seed=None):
    torch.manual_seed(seed)
    helio = helio_fn(N)
    field = ref_field.HelioField(helio, torch.tensor(target_pos), area, torch.tensor(target_normal),
                                 error_scale_mrad=err_mrad, sigma_scale=sigma_scale, resolution=R,
                                 device="cpu", max_batch_size=max(B, 1))
    if sun_fn is not None:
        sun = sun_fn(helio)
    elif single:
        sun = torch.tensor([700., 650., 720.])
    else:
        d = torch.nn.functional.normalize(torch.tensor([[0.5, 0.5, 0.7071]]) + 0.03 * torch.randn(B, 3), dim=1)
        sun = d * 14142.0
    ideal = field.calculate_ideal_normals(sun)
    field.init_actions(sun)
    action = field.initial_action.clone()
    action = action + 0.02 * torch.randn_like(action)       # not unit length on purpose (render does not normalise)
    if tweak is not None:
        action = tweak(action, field, sun)
    action = action.detach().requires_grad_(True)
    img, actual, refl = field.render(sun, action, ideal, monitor=True)
    # large cases: the image cotangent is regenerated from a frozen numpy stream instead of being stored
    # (np.random.RandomState is guaranteed stable across numpy versions; tests/conftest.py::load_golden rebuilds it)
    w_img = torch.randn_like(img) if w_seed is None else torch.from_numpy(
        np.random.RandomState(w_seed).standard_normal(tuple(img.shape)).astype(np.float32))
    w_act = torch.randn_like(actual)
    w_ref = torch.randn_like(refl)
    g_img_only, = torch.autograd.grad((img * w_img).sum(), action, retain_graph=True)
    loss = (img * w_img).sum() + (actual * w_act).sum() + (refl * w_ref).sum()
    g_all, = torch.autograd.grad(loss, action)
    errs = field.error_angles_mrad.unsqueeze(0) if (single or B == 1) else field.batch_error_angles_mrad[:B]
    np.savez_compressed(
        os.path.join(OUT, f"render_{name}.npz"),
        helio=npy(helio), target_pos=np.asarray(target_pos, np.float32), target_normal=np.asarray(target_normal, np.float32),
        area=np.asarray(area, np.float32), sigma_scale=np.float32(sigma_scale), err_mrad=np.float32(err_mrad),
        R=np.int32(R), single=np.bool_(single), sun=npy(sun), action=npy(action), errs=npy(errs), ideal=npy(ideal),
        error_angles_mrad=npy(field.error_angles_mrad), batch_error_angles_mrad=npy(field.batch_error_angles_mrad),
        img=npy(img), actual=npy(actual), refl=npy(refl),
        **(dict(w_img=npy(w_img)) if w_seed is None else dict(w_seed=np.int64(w_seed))), w_act=npy(w_act), w_ref=npy(w_ref),
        grad_img_only=npy(g_img_only), grad_all=npy(g_all), plane_u=npy(field.plane_u), plane_v=npy(field.plane_v),
        target_normal_unit=npy(field.target_normal))
    print(f"render_{name}: img max {float(img.max()):.4f} |grad| {float(g_all.abs().max()):.3e}")


def env_case(ref_env, name, seed, N, R, B, sigma_scale, err_mrad, helio_fn, **kw):
    torch.manual_seed(seed)
    helio = helio_fn(N)
    targ_pos = torch.tensor([0., -5., 0.])
    targ_norm = torch.tensor([0., 1., 0.])
    area = (15., 15.)
    env = ref_env.HelioEnv(heliostat_pos=helio, targ_pos=targ_pos, targ_area=area, targ_norm=targ_norm,
                           sigma_scale=sigma_scale, error_scale_mrad=err_mrad, initial_action_noise=0.0,
                           resolution=R, batch_size=B, device="cpu", new_sun_pos_every_reset=False,
                           new_errors_every_reset=True, **kw)
    sun_pos = env.sun_pos.clone()
    dmaps0 = env.distance_maps.clone()
    env.seed(seed + 1)
    obs0 = env.reset()
    reset_action = env.noisy_field.initial_action.clone()   # ideal + N(0, 0.01): HelioEnv never forwards its noise arg
    errs = env.noisy_field.batch_error_angles_mrad.clone()
    err1 = env.noisy_field.error_angles_mrad.clone()
    torch.manual_seed(seed + 2)
    action = (env.ideal_normals + 0.03 * torch.randn_like(env.ideal_normals)).flatten(1)
    action = action.detach().requires_grad_(True)
    obs, metrics, monitor = env.step(action)
    grads = {}
    for k in ("mse", "dist", "bound", "alignment_loss"):
        grads[k], = torch.autograd.grad(metrics[k], action, retain_graph=True, allow_unused=True)
    with torch.no_grad():
        target, _ = env.ref_field.render(env.sun_pos, env.ideal_normals.flatten(1), env.ideal_normals)
    np.savez_compressed(
        os.path.join(OUT, f"env_{name}.npz"),
        seed=np.int64(seed), helio=npy(helio), targ_pos=npy(targ_pos), targ_norm=npy(targ_norm), area=np.asarray(area, np.float32),
        sigma_scale=np.float32(sigma_scale), err_mrad=np.float32(err_mrad), R=np.int32(R), B=np.int32(B),
        sun_pos=npy(sun_pos), distance_maps=npy(dmaps0), ref_min=npy(env.ref_min), ref_max=npy(env.ref_max),
        errs=npy(errs), err_single=npy(err1), reset_img=npy(obs0["img"]), reset_aux=npy(obs0["aux"]), reset_action=npy(reset_action),
        ideal=npy(env.ideal_normals), action=npy(action), target=npy(target),
        ref_init_action=npy(env.ref_field.initial_action),   # ideal + N(0, 0.01) drawn inside set_sun_pos (test_environment.py:363)
        step_img=npy(obs["img"]), step_aux=npy(obs["aux"]),
        **{f"metric_{k}": npy(v) for k, v in metrics.items()},
        **{f"monitor_{k}": npy(v) for k, v in monitor.items()},
        **{f"grad_{k}": npy(v) for k, v in grads.items()},
        use_error_mask=np.bool_(kw.get("use_error_mask", False)), exponential_risk=np.bool_(kw.get("exponential_risk", False)))
    print(f"env_{name}: " + " ".join(f"{k}={float(v):.5g}" for k, v in metrics.items()))


def host_case(ref_env):
    """Seeded host-side helpers (sun cone sampling, az/el) for the not-gpu host-logic tests."""
    torch.manual_seed(7)
    axis = ref_env.azimuth_elevation_to_primary_direction(45.0, 45.0)
    dirs = ref_env.sample_cone_directions(9, axis, 2.0, force_upper_hemisphere=True)
    axis2 = ref_env.azimuth_elevation_to_primary_direction(10.0, 89.9)
    dirs2 = ref_env.sample_cone_directions(5, axis2, 2.0, force_upper_hemisphere=True)
    imgs = torch.rand(3, 12, 12)
    dm = ref_env.make_distance_maps(imgs)
    np.savez_compressed(os.path.join(OUT, "host_helpers.npz"), axis=npy(axis), dirs=npy(dirs), axis2=npy(axis2),
                        dirs2=npy(dirs2), imgs=npy(imgs), dmaps=npy(dm))


def com_case():
    """CenterOfMass2D (layers/center_of_mass.py) forward + autograd on seeded images, incl. a mass-free image,
    negative pixels (clamped), a (B,1,H,W) input and a non-square image."""
    sys.path.insert(0, REF)
    from layers.center_of_mass import CenterOfMass2D
    torch.manual_seed(31)
    com = CenterOfMass2D()
    out = {}
    for name, shape in (("sq", (5, 24, 24)), ("rect", (3, 1, 10, 37))):
        x = torch.rand(*shape)
        flat = x.view(shape[0], shape[-2], shape[-1])
        flat[1] = 0.0                                   # no mass -> (-1,-1), zero gradient
        flat[2] -= 0.5                                  # negative pixels are clamped to zero mass
        flat[0, 3, 4] = 0.0                             # clamp_min passes the gradient at exactly 0
        x = x.clone().requires_grad_(True)
        coords = com(x)
        w = torch.randn_like(coords)
        g, = torch.autograd.grad((coords * w).sum(), x)
        out.update({f"{name}_x": npy(x), f"{name}_coords": npy(coords), f"{name}_w": npy(w), f"{name}_grad": npy(g)})
    np.savez_compressed(os.path.join(OUT, "com.npz"), **out)
    print("com: ", out["sq_coords"][:3].tolist())


def angular_case(ref_field):
    """Angular action space (newenv/test_environment_angular.py:205-214): north-pointing normals rotated by the action's two
    angles with the reference's rotate_normals_batch, plus autograd of a weighted sum."""
    torch.manual_seed(41)
    B, N = 3, 7
    angles = (torch.randn(B, N * 2) * 400.0).requires_grad_(True)          # the helper scales by 1e-3: +-0.4 rad
    north = torch.zeros(B * N, 3)
    north[:, 1] = 1.0
    normals = ref_field.rotate_normals_batch(north, angles.view(-1, 2)).view(B, N, 3)
    w = torch.randn_like(normals)
    g, = torch.autograd.grad((normals * w).sum(), angles)
    np.savez_compressed(os.path.join(OUT, "angular.npz"), angles=npy(angles), normals=npy(normals), w=npy(w), grad=npy(g))
    print("angular:", normals[0, 0].tolist())


def main():
    os.makedirs(OUT, exist_ok=True)
    ref_field, ref_env = import_reference()
    # torch.linspace restatement check (oracle/helio_oracle.py::linspace_torch must be bit-exact)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helio_oracle import linspace_torch
    for w, n in [(15., 128), (15., 256), (15., 100), (7.3, 33), (12., 64), (15., 512), (15., 24)]:
        assert np.array_equal(torch.linspace(-w / 2, w / 2, n).numpy(), linspace_torch(-w / 2, w / 2, n)), (w, n)

    readme = lambda n: torch.cat([torch.rand(n, 2) * 10, torch.zeros(n, 1)], 1)            # README.md:69-70
    trainer = lambda n: torch.cat([torch.rand(n, 2) * 10 + 80, torch.zeros(n, 1)], 1)      # train_with_env.py:227

    def flip_one(action, field, sun):       # one mirror pointing below the horizon -> leaky-ReLU branch
        a = action.clone().view(action.shape[0], -1, 3)
        a[0, 1, 2] = -0.4
        a[-1, 0, 2] = -0.05
        return a.view(action.shape)

    render_case(ref_field, "readme", 11, N=6, R=24, B=4, helio_fn=readme, sigma_scale=0.1, err_mrad=90.0, tweak=flip_one)
    render_case(ref_field, "trainer", 12, N=7, R=32, B=3, helio_fn=trainer, sigma_scale=0.01, err_mrad=90.0)
    render_case(ref_field, "single", 13, N=5, R=16, B=1, helio_fn=readme, sigma_scale=0.1, err_mrad=30.0, single=True)
    render_case(ref_field, "tilted", 14, N=5, R=16, B=2, helio_fn=readme, sigma_scale=0.1, err_mrad=20.0,
                target_normal=(0.2, 1.0, 0.3), area=(12., 9.))
    render_case(ref_field, "wide", 15, N=9, R=40, B=2, helio_fn=trainer, sigma_scale=0.02, err_mrad=5.0, area=(15., 10.))

    # a ray exactly parallel to the receiver plane: valid=0 => +1 on every pixel (:63-75,141-143)
    def parallel_sun(helio):
        return torch.stack([helio[0] + torch.tensor([0., 0., 1000.]), helio[0] + torch.tensor([300., 200., 900.])])

    def parallel_action(action, field, sun):
        a = action.clone().view(2, -1, 3)
        a[0, 0] = torch.tensor([1.0, 0.0, 1.0])
        return a.view(action.shape)

    render_case(ref_field, "parallel", 16, N=3, R=16, B=2, helio_fn=readme, sigma_scale=0.1, err_mrad=0.0,
                tweak=parallel_action, sun_fn=parallel_sun)

    env_case(ref_env, "readme", 21, N=6, R=24, B=4, sigma_scale=0.1, err_mrad=90.0, helio_fn=readme)
    env_case(ref_env, "trainer", 22, N=8, R=32, B=5, sigma_scale=0.01, err_mrad=180.0, helio_fn=trainer)
    env_case(ref_env, "mask", 23, N=5, R=16, B=10, sigma_scale=0.1, err_mrad=90.0, helio_fn=readme, use_error_mask=True)
    env_case(ref_env, "exprisk", 24, N=5, R=16, B=3, sigma_scale=0.1, err_mrad=400.0, helio_fn=readme, exponential_risk=True)
    host_case(ref_env)
    com_case()
    angular_case(ref_field)

    # ---- shapes that route to the tcgen05 kernels (R >= 48): the tensor-core path pinned on the reference itself ----
    # BASELINE.json configs[0]/[1]: README quick start, N=50, 128x128, B=25, sigma_scale 0.1, 90 mrad
    render_case(ref_field, "c1", 31, N=50, R=128, B=25, helio_fn=readme, sigma_scale=0.1, err_mrad=90.0, tweak=flip_one, w_seed=131)
    # 256x256 receiver (BASELINE configs[3] resolution): NT=256 tiles, the CTA-pair (cta_group::2) kernels
    render_case(ref_field, "r256", 32, N=64, R=256, B=2, helio_fn=trainer, sigma_scale=0.01, err_mrad=90.0, w_seed=132)
    # 64x64: the NT=64 tile shape; N=130 spans two 128-heliostat blocks in the backward
    render_case(ref_field, "r64", 33, N=130, R=64, B=3, helio_fn=trainer, sigma_scale=0.02, err_mrad=30.0, w_seed=133)
    # BASELINE.json configs[1]: HelioEnv reset/step N=50, 128x128, B=25, new errors every reset, all four losses
    env_case(ref_env, "c2", 34, N=50, R=128, B=25, sigma_scale=0.1, err_mrad=90.0, helio_fn=readme)


if __name__ == "__main__":
    main()
