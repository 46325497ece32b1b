"""CPU oracle for the DOODLE flux-renderer hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain numpy restatement of the reference's algorithm for
``HelioField.render`` forward + backward and the ``HelioEnv.step`` loss block.
It exists so that the CUDA path can be checked on machines where
``/root/reference`` is absent.  Only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product package ``doodle_b200`` never does and has no CPU fallback.

Parity pinning: the reference ships no golden vectors, so the
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, generated in the build
container by ``oracle/make_golden.py`` (imports ``/root/reference``) and committed
under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks every function
here against those fixtures.

Every function cites the reference lines it follows (paths relative to the
reference root).  ``dtype=np.float32`` mirrors the reference's arithmetic op for
op; ``dtype=np.float64`` is the tie-breaker used when two fp32 results disagree
near tolerance.

The forward is the reference's DENSE algorithm (3-D differences per pixel, one
exp per heliostat-pixel) -- deliberately not the separable form the CUDA kernels
use, so that it is an independent check.  ``render_backward`` is the hand-written
adjoint of that dense forward (what autograd does in the reference).
"""
from __future__ import annotations

import math
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

LEAKY_SLOPE = 0.01  # F.leaky_relu default, newenv_rl_test_multi_error.py:369
_NCHUNK = 64         # heliostats per temporary block (memory bound only)


# ----------------------------------------------------------------------------
# small helpers
# ----------------------------------------------------------------------------
def _f(x, dtype):
    return np.asarray(x, dtype=dtype)


def _norm_last(x):
    """x.norm(dim=-1, keepdim=True)"""
    return np.sqrt((x * x).sum(axis=-1, keepdims=True))


def linspace_torch(start: float, end: float, steps: int, dtype=np.float32) -> np.ndarray:
    """torch.linspace as the CPU kernel evaluates it (newenv_rl_test_multi_error.py:129-130).

    step = (end-start)/(steps-1); the first half is fma(step, i, start), the second
    half fma(-step, steps-1-i, end).  Verified bit-exact against torch 2.11 CPU for
    fp32 in oracle/make_golden.py.
    """
    if steps == 1:
        return np.asarray([start], dtype=dtype)
    s = dtype(start)
    e = dtype(end)
    step = dtype((e - s) / dtype(steps - 1))
    i = np.arange(steps)
    half = steps // 2
    lo = np.float64(step) * i + np.float64(s)              # one rounding == fma
    hi = np.float64(e) - np.float64(step) * (steps - 1 - i)
    return np.where(i < half, lo, hi).astype(dtype)


def plane_basis(target_normal, dtype=np.float32):
    """Unit normal + in-plane basis (newenv_rl_test_multi_error.py:189-213)."""
    n = _f(target_normal, dtype)
    n = n / max(float(np.sqrt((n * n).sum())), 1e-9)
    n = n.astype(dtype)
    u = np.array([1.0, 0.0, 0.0], dtype=dtype)
    # torch.allclose(n, (0,1,0)): |a-b| <= 1e-8 + 1e-5*|b|
    ref = np.array([0.0, 1.0, 0.0], dtype=dtype)
    if np.all(np.abs(n - ref) <= 1e-8 + 1e-5 * np.abs(ref)):
        v = np.array([0.0, 0.0, 1.0], dtype=dtype)
    else:
        v = np.cross(n, u).astype(dtype)
        v = (v / max(float(np.sqrt((v * v).sum())), 1e-9)).astype(dtype)
    return n, u, v


# ----------------------------------------------------------------------------
# optics helpers
# ----------------------------------------------------------------------------
def calculate_ideal_normals(sun, helio, target_pos, dtype=np.float32):
    """newenv_rl_test_multi_error.py:256-278.  sun [B,3] -> [B,N,3]; sun [3] -> [N,3]."""
    sun = _f(sun, dtype)
    helio = _f(helio, dtype)
    tp = _f(target_pos, dtype)
    single = sun.ndim == 1
    s = sun.reshape(-1, 1, 3)
    inc = s - helio[None]
    refl = tp.reshape(1, 1, 3) - helio[None]
    inc_dir = inc / np.maximum(_norm_last(inc), dtype(1e-9))
    ref_dir = refl / np.maximum(_norm_last(refl), dtype(1e-9))
    nrm = inc_dir + ref_dir
    out = nrm / np.maximum(_norm_last(nrm), dtype(1e-9))
    out = out.astype(dtype)
    return out[0] if single else out


def rotate_normals(normals, errs_mrad):
    """rotate_normals_batch, newenv_rl_test_multi_error.py:78-104 ([M,3],[M,2])->[M,3]."""
    dtype = normals.dtype.type
    ae = errs_mrad[:, 0] * dtype(1e-3)
    au = errs_mrad[:, 1] * dtype(1e-3)
    ce, se = np.cos(ae), np.sin(ae)
    cu, su = np.cos(au), np.sin(au)
    x, y, z = normals[:, 0], normals[:, 1], normals[:, 2]
    x_u = cu * x - su * y
    y_u = su * x + cu * y
    y_e = ce * y_u - se * z
    z_e = se * y_u + ce * z
    return np.stack([x_u, y_e, z_e], axis=1).astype(normals.dtype)


def _geom_forward(sun, action, errs, helio, target_pos, n_hat, sigma_scale):
    """Per-(b,n) chain of render(), newenv_rl_test_multi_error.py:356-389 + :126-127.

    All inputs already of one dtype.  Returns a dict of [M,*] intermediates.
    """
    dtype = action.dtype.type
    B = sun.shape[0]
    N = helio.shape[0]
    M = B * N
    flats = action.reshape(M, 3)
    e = errs.reshape(M, 2)
    rot = rotate_normals(flats, e)                                  # :359
    zr = rot[:, 2]
    lz = np.where(zr > 0, zr, zr * dtype(LEAKY_SLOPE))               # :369
    v = np.stack([rot[:, 0], rot[:, 1], lz], axis=1)
    vn = np.maximum(_norm_last(v), dtype(1e-9))
    actual = v / vn                                                  # :372
    h = np.broadcast_to(helio[None], (B, N, 3)).reshape(M, 3)
    inc = (sun[:, None, :] - helio[None]).reshape(M, 3)              # :377
    inc_n = np.maximum(_norm_last(inc), dtype(1e-9))
    i_hat = inc / inc_n                                              # :380
    # reflect_vectors :46-50
    an = np.maximum(_norm_last(actual), dtype(1e-9))
    a_unit = actual / an
    dots = -(i_hat * a_unit).sum(axis=1, keepdims=True)
    r = -i_hat - dtype(2) * dots * a_unit
    rn = np.maximum(_norm_last(r), dtype(1e-9))
    r_hat = r / rn                                                   # :383
    # ray_plane_intersection_batch :52-75 (n_hat is already unit; reference renormalises)
    nn = n_hat / max(dtype(np.sqrt((n_hat * n_hat).sum())), dtype(1e-9))
    denom = (r_hat * nn).sum(axis=1, keepdims=True)
    valid = np.abs(denom) > dtype(1e-9)
    safe_denom = np.where(valid, denom, dtype(1e-9))
    num = ((target_pos[None] - h) * nn).sum(axis=1, keepdims=True)
    t = num / safe_denom
    safe_t = np.where(valid, t, dtype(0))
    inter = h + safe_t * r_hat
    P = np.where(valid, inter, dtype(0))
    vmask = valid.astype(action.dtype)                               # [M,1]
    dvec = P - h
    dist = np.sqrt((dvec * dvec).sum(axis=1))                        # :126
    sigma = np.maximum(dtype(sigma_scale) * dist, dtype(1e-9))       # :127
    two_s2 = np.maximum(dtype(2) * sigma * sigma, dtype(1e-12))      # :146
    return dict(rot=rot, v=v, vn=vn, actual=actual.astype(action.dtype), h=h, i_hat=i_hat,
                an=an, a_unit=a_unit, dots=dots, r=r, rn=rn, r_hat=r_hat.astype(action.dtype),
                nn=nn, denom=denom, valid=valid, num=num, t=safe_t, P=P.astype(action.dtype),
                vmask=vmask, dvec=dvec, dist=dist, sigma=sigma, two_s2=two_s2, errs=e)


def _grid_points(target_pos, u, v, width, height, R, dtype):
    """pts[R,R,3], newenv_rl_test_multi_error.py:129-138 (meshgrid 'ij': axis0 <-> x/u)."""
    xs = linspace_torch(-width / 2, width / 2, R, dtype)
    ys = linspace_torch(-height / 2, height / 2, R, dtype)
    gx, gy = np.meshgrid(xs, ys, indexing="ij")
    pts = (target_pos.reshape(1, 1, 3) + gx[..., None] * u.reshape(1, 1, 3)) + gy[..., None] * v.reshape(1, 1, 3)
    return pts.astype(dtype)


def _threads():
    return max(1, int(os.environ.get("HELIO_ORACLE_THREADS", os.cpu_count() or 1)))


def render_forward(sun, action, errs, helio, target_pos, target_normal, target_area,
                   resolution, sigma_scale, dtype=np.float32, keep=False, threads=None):
    """HelioField.render, newenv_rl_test_multi_error.py:308-415 (batched form).

    sun [B,3], action [B,N,3] (or [B,3N]), errs [B,N,2] (the tensor render() selected).
    Returns (images [B,R,R], actual [B,N,3], refl [B*N,3]) and, with keep=True, the
    intermediates needed by render_backward.
    """
    sun = _f(sun, dtype).reshape(-1, 3)
    helio = _f(helio, dtype)
    B, N = sun.shape[0], helio.shape[0]
    action = _f(action, dtype).reshape(B, N, 3)
    errs = _f(errs, dtype).reshape(B, N, 2)
    tp = _f(target_pos, dtype)
    n_hat, u, v = plane_basis(target_normal, dtype)
    width, height = float(target_area[0]), float(target_area[1])
    R = int(resolution)
    g = _geom_forward(sun, action, errs, helio, tp, n_hat, sigma_scale)
    pts = _grid_points(tp, u, v, width, height, R, dtype)            # [R,R,3]

    P = g["P"].reshape(B, N, 3)
    vm = g["vmask"].reshape(B, N, 1)
    ts2 = g["two_s2"].reshape(B, N)
    images = np.zeros((B, R, R), dtype=dtype)

    def one(b):
        # gaussian_blur_batch :140-148 for the N heliostats of sun b, then sum over N :406
        # (heliostats in chunks only to bound the [n,R,R,3] temporaries; the arithmetic is unchanged)
        acc = np.zeros((R, R), dtype=dtype)
        for n0 in range(0, N, _NCHUNK):
            s = slice(n0, min(N, n0 + _NCHUNK))
            diffs = (pts[None] - P[b][s, None, None, :]) * vm[b][s, None, None, :]
            dist_sq = (diffs * diffs).sum(axis=3)
            G = np.exp(-dist_sq / ts2[b][s, None, None])
            acc += G.sum(axis=0)
        images[b] = acc

    th = threads or _threads()
    if th > 1 and B > 1:
        with ThreadPoolExecutor(th) as ex:
            list(ex.map(one, range(B)))
    else:
        for b in range(B):
            one(b)
    out = (images, g["actual"].reshape(B, N, 3), g["r_hat"])
    if keep:
        g.update(pts=pts, B=B, N=N, R=R, sigma_scale=sigma_scale)
        return out, g
    return out


def render_backward(ctx, g_img=None, g_actual=None, g_refl=None, threads=None):
    """Adjoint of render_forward w.r.t. ``action`` (what autograd does in the reference).

    ctx is the dict returned by render_forward(keep=True).  Returns dL/daction [B,N,3].
    """
    g = ctx
    B, N, R = g["B"], g["N"], g["R"]
    M = B * N
    arr = g["actual"]
    dtype = arr.dtype.type
    z = lambda *s: np.zeros(s, dtype=arr.dtype)
    gP = z(M, 3)
    g_ts2 = z(M)
    if g_img is not None:
        g_img = _f(g_img, arr.dtype).reshape(B, R, R)
        P = g["P"].reshape(B, N, 3)
        vm = g["vmask"].reshape(B, N, 1)
        ts2 = g["two_s2"].reshape(B, N)
        pts = g["pts"]
        gP3 = gP.reshape(B, N, 3)
        gts = g_ts2.reshape(B, N)

        def one(b):
            for n0 in range(0, N, _NCHUNK):
                s = slice(n0, min(N, n0 + _NCHUNK))
                diffs = (pts[None] - P[b][s, None, None, :]) * vm[b][s, None, None, :]
                dist_sq = (diffs * diffs).sum(axis=3)
                G = np.exp(-dist_sq / ts2[b][s, None, None])
                gG = g_img[b][None] * G                               # dL/d(exp arg)
                g_dsq = -gG / ts2[b][s, None, None]
                gts[b, s] = (gG * dist_sq).sum(axis=(1, 2)) / (ts2[b][s] * ts2[b][s])
                g_diffs = dtype(2) * diffs * g_dsq[..., None] * vm[b][s, None, None, :]
                gP3[b, s] = -g_diffs.sum(axis=(1, 2))

        th = threads or _threads()
        if th > 1 and B > 1:
            with ThreadPoolExecutor(th) as ex:
                list(ex.map(one, range(B)))
        else:
            for b in range(B):
                one(b)
    # two_s2 = clamp_min(2 sigma^2, 1e-12); sigma = clamp_min(scale*dist, 1e-9); dist = |P-h|
    sigma, dist = g["sigma"], g["dist"]
    ss = dtype(g["sigma_scale"])
    live = (dtype(2) * sigma * sigma) >= dtype(1e-12)
    g_sigma = np.where(live, g_ts2 * dtype(4) * sigma, dtype(0))
    live2 = (ss * dist) >= dtype(1e-9)
    g_dist = np.where(live2, g_sigma * ss, dtype(0))
    safe_dist = np.where(dist > 0, dist, dtype(1))
    gP = gP + np.where(dist[:, None] > 0, g_dist[:, None] * g["dvec"] / safe_dist[:, None], dtype(0))
    # P = where(valid, h + t*r_hat, 0)
    valid = g["valid"]
    gP = np.where(valid, gP, dtype(0))
    g_rhat = g["t"] * gP
    if g_refl is not None:
        g_rhat = g_rhat + _f(g_refl, arr.dtype).reshape(M, 3)
    g_t = (gP * g["r_hat"]).sum(axis=1, keepdims=True)
    g_t = np.where(valid, g_t, dtype(0))
    # t = num / safe_denom ; safe_denom = where(valid, denom, eps)
    safe_denom = np.where(valid, g["denom"], dtype(1e-9))
    g_denom = np.where(valid, -g_t * g["num"] / (safe_denom * safe_denom), dtype(0))
    g_rhat = g_rhat + g_denom * g["nn"][None]
    # r_hat = r / clamp_min(|r|, 1e-9)
    rn, r_hat = g["rn"], g["r_hat"]
    g_r = (g_rhat - r_hat * (g_rhat * r_hat).sum(axis=1, keepdims=True)) / rn
    # r = -i - 2*dots*a_unit ; dots = -(i . a_unit)   (i carries no grad)
    i_hat, a_unit, dots = g["i_hat"], g["a_unit"], g["dots"]
    g_dots = -dtype(2) * (g_r * a_unit).sum(axis=1, keepdims=True)
    g_aunit = -dtype(2) * dots * g_r - g_dots * i_hat
    # a_unit = actual / clamp_min(|actual|, 1e-9)
    an = g["an"]
    g_act = (g_aunit - a_unit * (g_aunit * a_unit).sum(axis=1, keepdims=True)) / an
    if g_actual is not None:
        g_act = g_act + _f(g_actual, arr.dtype).reshape(M, 3)
    # actual = v / clamp_min(|v|, 1e-9)
    vn, actual = g["vn"], g["actual"]
    g_v = (g_act - actual * (g_act * actual).sum(axis=1, keepdims=True)) / vn
    # v_z = leaky_relu(rot_z)
    zr = g["rot"][:, 2]
    g_rot = g_v.copy()
    g_rot[:, 2] = g_v[:, 2] * np.where(zr > 0, dtype(1), dtype(LEAKY_SLOPE))
    # rot = Rx(e0) Rz(e1) n  -> transpose
    e = g["errs"]
    ae = e[:, 0] * dtype(1e-3)
    au = e[:, 1] * dtype(1e-3)
    ce, se, cu, su = np.cos(ae), np.sin(ae), np.cos(au), np.sin(au)
    gx_u = g_rot[:, 0]
    gy_u = ce * g_rot[:, 1] + se * g_rot[:, 2]
    gz = -se * g_rot[:, 1] + ce * g_rot[:, 2]
    gx = cu * gx_u + su * gy_u
    gy = -su * gx_u + cu * gy_u
    return np.stack([gx, gy, gz], axis=1).reshape(B, N, 3).astype(arr.dtype)


# ----------------------------------------------------------------------------
# HelioEnv.step pieces
# ----------------------------------------------------------------------------
def boundary(vects, helio, targ_pos, targ_norm, targ_area, u, v, dtype=np.float32, with_grad=False):
    """boundary(..., return_all=True), test_environment.py:101-130.  vects [B,N,3] -> [B,N].

    with_grad=True also returns d out[b,n] / d vects[b,n,:]  ([B,N,3]).
    """
    vects = _f(vects, dtype)
    helio = _f(helio, dtype)
    tp = _f(targ_pos, dtype)
    tn = _f(targ_norm, dtype)
    u = _f(u, dtype)
    v = _f(v, dtype)
    tol = dtype(0.75)
    dots = (-vects * tn).sum(-1)
    eps = dtype(1e-6)
    valid = np.abs(dots) > eps
    den = dots + (~valid).astype(dtype) * eps
    pv = (vects * tp).sum(-1)
    t = pv / den
    inter = helio[None] + vects * t[..., None]
    local = inter - tp
    xl = (local * u).sum(-1)
    yl = (local * v).sum(-1)
    hw = dtype(targ_area[0]) * tol / dtype(2)
    hh = dtype(targ_area[1]) * tol / dtype(2)
    ax, ay = np.abs(xl) - hw * tol, np.abs(yl) - hh * tol
    dx, dy = np.maximum(ax, 0), np.maximum(ay, 0)
    dist = np.sqrt(dx * dx + dy * dy + dtype(1e-8))
    inside = (np.abs(xl) <= hw) & (np.abs(yl) <= hh) & valid
    outm = (~inside).astype(dtype)
    out = (dist * outm).astype(dtype)
    if not with_grad:
        return out
    g_dx = outm * dx / dist * (ax > 0)
    g_dy = outm * dy / dist * (ay > 0)
    g_xl = g_dx * np.sign(xl)
    g_yl = g_dy * np.sign(yl)
    g_inter = g_xl[..., None] * u + g_yl[..., None] * v
    g_vec = g_inter * t[..., None]
    g_t = (g_inter * vects).sum(-1)
    g_pv = g_t / den
    g_den = -g_t * pv / (den * den)
    g_vec = g_vec + g_pv[..., None] * tp - g_den[..., None] * tn
    return out, g_vec.astype(dtype)


def angles_mrad(v1, v2, dtype=np.float32, with_grad=False):
    """calculate_angles_mrad, test_environment.py:132-155 (the 1e-10 epsilon vanishes in fp32
    but not in fp64 -- kept as written).  with_grad returns d angle / d v2."""
    v1 = _f(v1, dtype)
    v2 = _f(v2, dtype)
    dot = (v1 * v2).sum(-1)
    one = dtype(1.0)
    upper = np.nextafter(one, dtype(0.0))
    lo = dtype(float(-upper) + 1e-10)
    hi = dtype(float(upper) - 1e-10)
    c = np.clip(dot, lo, hi)
    ang = (np.arccos(c) * dtype(1000)).astype(dtype)
    if not with_grad:
        return ang
    passes = (dot >= lo) & (dot <= hi)
    g_dot = np.where(passes, -dtype(1000) / np.sqrt(one - c * c), dtype(0))
    return ang, (g_dot[..., None] * v1).astype(dtype)


def loss_block(img, target, dmaps, dtype=np.float32, with_grad=False, w_mse=1.0, w_dist=1.0):
    """HelioEnv.step image losses (use_error_mask=False), test_environment.py:436-457,492.

    Returns dict(mse, dist, mae_image[B,1], tx[B]); with_grad adds
    g_img = d(w_mse*mse + w_dist*dist)/d img.
    """
    img = _f(img, dtype)
    target = _f(target, dtype)
    dmaps = _f(dmaps, dtype)
    B = img.shape[0]
    tx = np.maximum(target.max(axis=(1, 2), keepdims=True), dtype(1e-6))
    pred_n = img / tx
    targ_n = target / tx
    diff = pred_n - targ_n
    err = np.abs(diff)
    mse = (diff * diff).mean(dtype=dtype)
    dist = (err * dmaps).sum(axis=(1, 2)).mean()
    mae = err.mean(axis=(1, 2)).reshape(-1, 1)
    out = dict(mse=dtype(mse), dist=dtype(dist), mae_image=mae.astype(dtype), tx=tx.reshape(-1))
    if with_grad:
        n_el = dtype(img.size)
        g = dtype(w_mse) * dtype(2) * diff / (n_el * tx) + dtype(w_dist) * np.sign(diff) * dmaps / (dtype(B) * tx)
        out["g_img"] = g.astype(dtype)
    return out


def env_step(sun, action, errs, helio, targ_pos, targ_norm, targ_area, resolution, sigma_scale,
             dmaps, weights=(1.0, 1.0, 1.0, 1.0), dtype=np.float32, target=None, threads=None):
    """HelioEnv.step forward + backward of sum_k w_k * metric_k w.r.t. action
    (test_environment.py:402-516, use_error_mask=False, exponential_risk=False).

    weights = (w_mse, w_dist, w_bound, w_alignment).  Returns (metrics dict, monitor dict,
    grad [B,N,3], img [B,R,R]).  ``target`` may be passed to skip the second render.
    """
    sun = _f(sun, dtype).reshape(-1, 3)
    helio = _f(helio, dtype)
    B, N = sun.shape[0], helio.shape[0]
    action = _f(action, dtype).reshape(B, N, 3)
    ideal = calculate_ideal_normals(sun, helio, targ_pos, dtype)                   # :414
    (img, actual, refl), ctx = render_forward(sun, action, errs, helio, targ_pos, targ_norm, targ_area,
                                              resolution, sigma_scale, dtype, keep=True, threads=threads)
    if target is None:                                                             # :429-435
        target, _, _ = render_forward(sun, ideal, np.zeros((B, N, 2), dtype), helio, targ_pos, targ_norm,
                                      targ_area, resolution, sigma_scale, dtype, threads=threads)
    w_mse, w_dist, w_bound, w_align = weights
    lb = loss_block(img, target, dmaps, dtype, with_grad=True, w_mse=w_mse, w_dist=w_dist)
    ang, g_ang = angles_mrad(ideal, actual, dtype, with_grad=True)                 # :455
    u = np.array([1.0, 0.0, 0.0], dtype)
    v = np.array([0.0, 0.0, 1.0], dtype)
    bnd, g_bnd = boundary(action, helio, targ_pos, targ_norm, targ_area, u, v, dtype, with_grad=True)
    metrics = dict(mse=lb["mse"], dist=lb["dist"], bound=bnd.mean(dtype=dtype), alignment_loss=ang.mean(dtype=dtype))
    g_actual = g_ang * dtype(w_align) / dtype(B * N)
    grad = render_backward(ctx, g_img=lb["g_img"], g_actual=g_actual, threads=threads)
    grad = grad + g_bnd * dtype(w_bound) / dtype(B * N)
    monitor = dict(normals=action, reflected_rays=refl, ideal_normals=ideal.reshape(-1, 3), all_bounds=bnd,
                   mae_image=lb["mae_image"], alignment_errors=ang.reshape(-1), target=target, actual=actual)
    return metrics, monitor, grad.astype(dtype), img


# ----------------------------------------------------------------------------
# sun sampling (setup-time host logic the env shim mirrors)
# ----------------------------------------------------------------------------
def azimuth_elevation_to_direction(az_deg: float, el_deg: float) -> np.ndarray:
    """test_environment.py:18-40"""
    az, el = math.radians(az_deg), math.radians(el_deg)
    vec = np.array([math.cos(el) * math.cos(az), math.cos(el) * math.sin(az), math.sin(el)], dtype=np.float32)
    return vec / np.sqrt((vec * vec).sum(dtype=np.float32))


# ----------------------------------------------------------------------------
# centre of mass (the COM trainer's encoder feed)
# ----------------------------------------------------------------------------
def center_of_mass(x, eps=1e-12, g_coords=None, dtype=np.float32):
    """layers/center_of_mass.py:21-60 (CenterOfMass2D.forward) and its autograd.

    x [B,H,W] -> coords [B,2] = (x_com, y_com), origin top-left, x = column, y = row; (-1,-1) when the
    image has no mass.  With ``g_coords`` [B,2] also returns dL/dx [B,H,W] (clamp_min passes the gradient
    where x >= 0; the (-1,-1) overwrite kills it for mass-free images)."""
    x = _f(x, dtype)
    B, H, W = x.shape
    w = np.maximum(x, dtype(0))                                                    # :37
    yy, xx = np.meshgrid(np.arange(H, dtype=dtype), np.arange(W, dtype=dtype), indexing="ij")   # :40-44
    w_sum = w.sum(axis=(1, 2), dtype=dtype)                                        # :47
    x_wsum = (w * xx).sum(axis=(1, 2), dtype=dtype)
    y_wsum = (w * yy).sum(axis=(1, 2), dtype=dtype)
    den = w_sum + dtype(eps)
    coords = np.stack([x_wsum / den, y_wsum / den], axis=-1).astype(dtype)         # :52-55
    no_mass = w_sum <= 0                                                           # :58-60
    coords[no_mass] = dtype(-1.0)
    if g_coords is None:
        return coords
    g = _f(g_coords, dtype)
    gx = np.where(no_mass, dtype(0), g[:, 0] / den)[:, None, None]
    gy = np.where(no_mass, dtype(0), g[:, 1] / den)[:, None, None]
    xc = (x_wsum / den)[:, None, None]
    yc = (y_wsum / den)[:, None, None]
    g_x = (gx * (xx[None] - xc) + gy * (yy[None] - yc)) * (x >= 0)
    return coords, g_x.astype(dtype)
