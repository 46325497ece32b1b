"""Pin the numpy oracle against outputs of the reference itself (tests/golden, made by
oracle/make_golden.py).  CPU only.  Tolerances: images rtol 1e-4 / atol 1e-6 and action
gradients 1e-3 relative (BASELINE.json north_star)."""
import numpy as np
import pytest

from conftest import load_golden, rel_err
from oracle import helio_oracle as orc

RENDER = ["readme", "trainer", "single", "tilted", "wide", "parallel"]
ENV = ["readme", "trainer", "exprisk"]


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
        np.testing.assert_allclose(ideal, g["ideal"], rtol=1e-6, atol=1e-7)


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
    np.testing.assert_allclose(monitor["alignment_errors"], g["monitor_alignment_errors"], rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(monitor["mae_image"], g["monitor_mae_image"], rtol=2e-4, atol=1e-7)
    np.testing.assert_allclose(monitor["reflected_rays"], g["monitor_reflected_rays"], rtol=1e-5, atol=1e-6)
    for k, w in (("mse", (1, 0, 0, 0)), ("dist", (0, 1, 0, 0)), ("bound", (0, 0, 1, 0)), ("alignment_loss", (0, 0, 0, 1))):
        if k == "bound" and bool(g["exponential_risk"]):
            continue
        _, _, grad, _ = orc.env_step(*args, weights=w, target=g["target"])
        ref = g["grad_" + k].reshape(B, -1, 3)
        assert rel_err(grad, ref) < 1e-3, k


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
