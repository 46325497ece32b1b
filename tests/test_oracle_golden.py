"""Pin the numpy oracle against outputs of the reference itself (tests/golden, made by
oracle/make_golden.py).  CPU only.  Tolerances: images rtol 1e-4 / atol 1e-6 and action
gradients 1e-3 relative (BASELINE.json north_star)."""
import numpy as np
import pytest

from conftest import acos_grad_slack, assert_grad_close, load_golden, rel_err
from oracle import helio_oracle as orc

RENDER = ["readme", "trainer", "single", "tilted", "wide", "parallel", "c1", "r256", "r64"]   # c1 / r256 / r64: tensor-core shapes
ENV = ["readme", "trainer", "exprisk", "c2"]


def _render(g, dtype):
    sun = g["sun"].reshape(-1, 3)
    return orc.render_forward(sun, g["action"], g["errs"], g["helio"], g["target_pos"], g["target_normal"],
                              tuple(g["area"]), int(g["R"]), float(g["sigma_scale"]), dtype=dtype, keep=True)


@pytest.mark.parametrize("name", RENDER)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_render_forward_matches_reference(name, dtype):
    g = load_golden("render_" + name)
    (img, actual, refl), _ = _render(g, dtype)
    ref_img = g["img"].reshape(img.shape)
    np.testing.assert_allclose(img, ref_img, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(actual.reshape(-1, 3), g["actual"].reshape(-1, 3), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(refl, g["refl"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", RENDER)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_render_backward_matches_reference_autograd(name, dtype):
    g = load_golden("render_" + name)
    (img, actual, refl), ctx = _render(g, dtype)
    gi = orc.render_backward(ctx, g_img=g["w_img"])
    ga = orc.render_backward(ctx, g_img=g["w_img"], g_actual=g["w_act"], g_refl=g["w_ref"])
    assert rel_err(gi.reshape(-1), g["grad_img_only"].reshape(-1)) < 1e-3
    assert rel_err(ga.reshape(-1), g["grad_all"].reshape(-1)) < 1e-3
    if dtype == np.float64:     # elementwise: the fp64 adjoint against the reference's fp32 autograd, component by component
        assert_grad_close(g["grad_img_only"].reshape(-1, 3), gi.reshape(-1, 3), what=f"{name} img-only (reference vs fp64 oracle)")
        assert_grad_close(g["grad_all"].reshape(-1, 3), ga.reshape(-1, 3), what=f"{name} all (reference vs fp64 oracle)")


def test_parallel_ray_adds_one_everywhere():
    g = load_golden("render_parallel")
    (img, _, _), ctx = _render(g, np.float32)
    assert not ctx["valid"].reshape(2, -1)[0, 0] and ctx["valid"].sum() == ctx["valid"].size - 1
    assert img[0].min() >= 1.0 - 1e-6            # G == 1 for the invalid ray (newenv_rl_test_multi_error.py:141-143)
    np.testing.assert_allclose(img, g["img"], rtol=1e-4, atol=1e-6)


def test_ideal_normals():
    for name in RENDER:
        g = load_golden("render_" + name)
        ideal = orc.calculate_ideal_normals(g["sun"], g["helio"], g["target_pos"])
        np.testing.assert_allclose(ideal, g["ideal"], rtol=1e-6, atol=2.5e-7)     # unit vectors: 2 ulp of 1.0


def test_plane_basis():
    for name in RENDER:
        g = load_golden("render_" + name)
        n, u, v = orc.plane_basis(g["target_normal"])
        np.testing.assert_allclose(n, g["target_normal_unit"], atol=1e-7)
        np.testing.assert_allclose(u, g["plane_u"], atol=1e-7)
        np.testing.assert_allclose(v, g["plane_v"], atol=1e-7)


@pytest.mark.parametrize("name", ENV)
def test_env_step_matches_reference(name):
    g = load_golden("env_" + name)
    B = int(g["B"])
    args = (g["sun_pos"], g["action"], g["errs"], g["helio"], g["targ_pos"], g["targ_norm"], tuple(g["area"]),
            int(g["R"]), float(g["sigma_scale"]), g["distance_maps"])
    metrics, monitor, _, img = orc.env_step(*args, weights=(0, 0, 0, 0))
    np.testing.assert_allclose(img, g["step_img"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(monitor["target"], g["target"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(metrics["mse"], g["metric_mse"], rtol=2e-4)
    np.testing.assert_allclose(metrics["dist"], g["metric_dist"], rtol=2e-4)
    np.testing.assert_allclose(metrics["alignment_loss"], g["metric_alignment_loss"], rtol=1e-4)
    if not bool(g["exponential_risk"]):
        np.testing.assert_allclose(metrics["bound"], g["metric_bound"], rtol=1e-4, atol=1e-6)
    else:   # test_environment.py:472-480
        np.testing.assert_allclose(np.exp(monitor["all_bounds"] + np.float32(1e-6)).mean(), g["metric_bound"], rtol=1e-4)
    np.testing.assert_allclose(monitor["all_bounds"], g["monitor_all_bounds"], rtol=1e-4, atol=1e-5)
    # angle = 1000 acos(dot): a 2-ulp difference of the fp32 dot product (~2.4e-7 near 1) moves the angle by
    # 1000 * 2.4e-7 / sin(angle) mrad, which dominates for well-aligned mirrors (floor: the acos clamp, 0.345 mrad)
    ref_a = g["monitor_alignment_errors"]
    tol_a = 1e-4 * np.abs(ref_a) + 1e-3 + 1000.0 * 2.4e-7 / np.maximum(np.sin(ref_a * 1e-3), 3.4e-4)
    assert np.all(np.abs(monitor["alignment_errors"] - ref_a) <= tol_a), float(np.abs(monitor["alignment_errors"] - ref_a).max())
    np.testing.assert_allclose(monitor["mae_image"], g["monitor_mae_image"], rtol=2e-4, atol=1e-7)
    np.testing.assert_allclose(monitor["reflected_rays"], g["monitor_reflected_rays"], rtol=1e-5, atol=1e-6)
    for k, w in (("mse", (1, 0, 0, 0)), ("dist", (0, 1, 0, 0)), ("bound", (0, 0, 1, 0)), ("alignment_loss", (0, 0, 0, 1))):
        if k == "bound" and bool(g["exponential_risk"]):
            continue
        _, _, grad, _ = orc.env_step(*args, weights=w, target=g["target"])
        ref = g["grad_" + k].reshape(B, -1, 3)
        slack = acos_grad_slack(ref_a.reshape(B, -1)) if k == "alignment_loss" else None
        if slack is None:
            assert rel_err(grad, ref) < 1e-3, k
        assert_grad_close(grad, ref, what=f"env_{name} d{k}/daction (fp32 oracle vs reference autograd)", extra_rtol=slack)


def test_env_reset_image_and_aux():
    g = load_golden("env_readme")
    img, _, _ = orc.render_forward(g["sun_pos"], g["reset_action"], g["errs"], g["helio"], g["targ_pos"], g["targ_norm"],
                                   tuple(g["area"]), int(g["R"]), float(g["sigma_scale"]))
    np.testing.assert_allclose(img, g["reset_img"], rtol=1e-4, atol=1e-6)
    aux = np.concatenate([g["sun_pos"], g["ideal"].reshape(len(g["sun_pos"]), -1)], 1)
    np.testing.assert_allclose(aux, g["reset_aux"], rtol=1e-6)


def test_fp64_and_fp32_oracles_agree():
    g = load_golden("render_trainer")
    (i32, _, _), c32 = _render(g, np.float32)
    (i64, _, _), c64 = _render(g, np.float64)
    np.testing.assert_allclose(i32, i64, rtol=1e-4, atol=1e-6)
    g32 = orc.render_backward(c32, g_img=g["w_img"])
    g64 = orc.render_backward(c64, g_img=g["w_img"])
    assert rel_err(g32, g64) < 1e-3


@pytest.mark.parametrize("name", ["sq", "rect"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_center_of_mass_matches_reference(name, dtype):
    """oracle.center_of_mass vs the reference's CenterOfMass2D forward + autograd (tests/golden/com.npz)."""
    g = load_golden("com")
    x = g[name + "_x"]
    x3 = x.reshape(x.shape[0], x.shape[-2], x.shape[-1])
    coords, grad = orc.center_of_mass(x3, g_coords=g[name + "_w"], dtype=dtype)
    np.testing.assert_allclose(coords, g[name + "_coords"], rtol=2e-6, atol=1e-6)
    assert np.array_equal(coords[1], [-1.0, -1.0])
    assert rel_err(grad.reshape(-1), g[name + "_grad"].reshape(-1)) < 1e-5
    assert not grad[1].any()                                   # mass-free image: zero gradient


# ---- the torch-eager dense restatement (bench.py's same-box GPU comparator) pinned on the same fixtures ----------------
@pytest.mark.parametrize("name", ["readme", "trainer", "tilted", "parallel", "r64"])
def test_torch_eager_render_matches_reference(name):
    import torch
    from oracle import helio_torch_eager as te
    g = load_golden("render_" + name)
    t = lambda k: torch.as_tensor(g[k])
    sun = t("sun").reshape(-1, 3)
    B = sun.shape[0]
    action = t("action").reshape(B, -1).clone().requires_grad_(True)
    img, actual, refl = te.render(sun, action, t("errs"), t("helio"), t("target_pos"), t("target_normal"), tuple(float(x) for x in g["area"]),
                                  int(g["R"]), float(g["sigma_scale"]))
    np.testing.assert_allclose(img.detach().numpy(), g["img"].reshape(img.shape), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(actual.detach().numpy(), g["actual"].reshape(actual.shape), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(refl.detach().numpy(), g["refl"], rtol=1e-5, atol=1e-6)
    loss = (img * t("w_img").reshape(img.shape)).sum() + (actual * t("w_act").reshape(actual.shape)).sum() + (refl * t("w_ref")).sum()
    ga, = torch.autograd.grad(loss, action)
    assert_grad_close(ga.numpy().reshape(-1, 3), g["grad_all"].reshape(-1, 3), what=f"torch-eager render_{name}")
    np.testing.assert_allclose(te.ideal_normals(sun, t("helio"), t("target_pos")).numpy(), g["ideal"].reshape(B, -1, 3), rtol=1e-6, atol=2.5e-7)


@pytest.mark.parametrize("name", ["readme", "trainer"])
def test_torch_eager_env_step_matches_reference(name):
    import torch
    from oracle import helio_torch_eager as te
    g = load_golden("env_" + name)
    t = lambda k: torch.as_tensor(g[k])
    B = int(g["B"])
    action = t("action").clone().requires_grad_(True)
    m, img, target = te.env_step(t("sun_pos"), action, t("errs"), t("helio"), t("targ_pos"), t("targ_norm"), tuple(float(x) for x in g["area"]),
                                 int(g["R"]), float(g["sigma_scale"]), t("distance_maps"))
    np.testing.assert_allclose(img.detach().numpy(), g["step_img"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(target.numpy(), g["target"], rtol=1e-4, atol=1e-6)
    for k in ("mse", "dist", "bound", "alignment_loss"):
        np.testing.assert_allclose(float(m[k]), float(g["metric_" + k]), rtol=2e-4, err_msg=k)
        gr, = torch.autograd.grad(m[k], action, retain_graph=True)
        slack = acos_grad_slack(g["monitor_alignment_errors"].reshape(B, -1)) if k == "alignment_loss" else None
        assert_grad_close(gr.numpy().reshape(B, -1, 3), g["grad_" + k].reshape(B, -1, 3), what=f"torch-eager env_{name} {k}", extra_rtol=slack)
    # the chunked driver bench.py times: same gradient as one un-chunked backward of the summed losses
    full, = torch.autograd.grad(m["mse"] + m["dist"] + m["bound"] + m["alignment_loss"], action)
    ch = te.chunked_step_and_backward(t("sun_pos"), t("action").view(B, -1, 3), t("errs"), t("helio"), t("targ_pos"), t("targ_norm"),
                                      tuple(float(x) for x in g["area"]), int(g["R"]), float(g["sigma_scale"]), t("distance_maps"), chunk=2)
    assert rel_err(ch.numpy().reshape(-1), full.numpy().reshape(-1)) < 1e-5
