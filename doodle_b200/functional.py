"""torch.autograd.Function wrappers over the C ABI (include/helio_b200.h).

PyTorch is plumbing here: it owns device memory, the current stream and the autograd tape; every
number is produced by the kernels in libhelio_sm100.so.  Nothing in this module has a CPU path.

Graph for one differentiable render (HelioField.render, newenv_rl_test_multi_error.py:308-415):

    action --GeomFn--> params --SplatFn--> img
                   \\-> actual, refl, bounds, angles, sums

SplatFn.backward hands the per-(b,n) moments {S0,Sx,Sy,S2} of g*G to GeomFn.backward as the
"gradient" of ``params``; GeomFn.backward turns them into dL/daction (see include/helio_b200.h).
``params`` is an internal tensor, never exposed to callers.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from typing import Optional

import torch

from . import _lib
from ._lib import SPLAT_AUTO, SPLAT_SIMT, SPLAT_TC, Feed, Scene  # noqa: F401


# ---- launch accounting / live per-kernel timing (bench.py reads these) ------------------------
_LAUNCHES = 0


def launch_count() -> int:
    """Number of libhelio kernels launched by this process so far."""
    return _LAUNCHES


_PROFILE_ON = False


def profiling() -> bool:
    """True while helio_profile_enable(1) is in effect (event pairs cannot be recorded into a CUDA graph: HelioEnv then
    keeps to the eager step)."""
    return _PROFILE_ON


def reset_profile(enabled: bool):
    """Switch the library's per-kernel CUDA-event timing on/off (helio_profile_enable); clears earlier records."""
    global _PROFILE_ON
    _lib.check(_lib.load().helio_profile_enable(1 if enabled else 0), "helio_profile_enable")
    _PROFILE_ON = bool(enabled)


def collect_profile():
    """{kernel: {n, total_ms, avg_ms}} from the CUDA events the library recorded around each kernel on its stream."""
    lib = _lib.load()
    out = {}
    name, ms = C.c_char_p(), C.c_float()
    for i in range(lib.helio_profile_count()):
        _lib.check(lib.helio_profile_get(i, C.byref(name), C.byref(ms)), "helio_profile_get")
        d = out.setdefault(name.value.decode(), dict(n=0, total_ms=0.0))
        d["n"] += 1
        d["total_ms"] += ms.value
    for d in out.values():
        d["avg_ms"] = d["total_ms"] / d["n"]
    return out


def tc_clock_mhz():
    """(forward, backward) SM clock in MHz held inside the most recent tcgen05 splat kernels (helio_tc_clock_mhz; syncs)."""
    lib = _lib.load()
    out = []
    for which in (0, 1):
        v = C.c_float()
        _lib.check(lib.helio_tc_clock_mhz(which, C.byref(v)), "helio_tc_clock_mhz")
        out.append(float(v.value))
    return tuple(out)


class _Call:
    """Context manager around one C-ABI launch: device guard (only when the tensors live on another
    device than the current one) and launch count."""

    __slots__ = ("guard",)

    def __init__(self, name: str, device: torch.device):
        self.guard = None if device.index is None or device.index == torch._C._cuda_getDevice() else torch.cuda.device(device)

    def __enter__(self):
        global _LAUNCHES
        if self.guard is not None:
            self.guard.__enter__()
        _LAUNCHES += 1
        return self

    def __exit__(self, *exc):
        if self.guard is not None:
            return self.guard.__exit__(*exc)
        return False


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()      # ctypes converts int -> c_void_p for declared argtypes


def _stream():
    """cudaStream_t of torch's current stream on the current device (raw handle: the Python Stream object costs
    ~20 us to build, which is most of a small field's step)."""
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def _cf(t: torch.Tensor) -> torch.Tensor:
    """contiguous fp32 (mirrors torch.as_tensor(..., float32), newenv_rl_test_multi_error.py:326,334)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def require_cuda(device: torch.device, what: str):
    if device.type != "cuda":
        raise RuntimeError(f"{what}: doodle_b200 runs on sm_100a GPUs only and has no CPU fallback (got device={device})")
    if not torch.cuda.is_available():
        raise RuntimeError(f"{what}: CUDA is not available; doodle_b200 has no CPU fallback")


class GeomFn(torch.autograd.Function):
    """K1: action -> (params, actual, refl[, ideal, bounds, angles, sums])."""

    @staticmethod
    def forward(ctx, action, sun, errs, helio, scene, workspace, want_aux: bool):
        lib = _lib.load()
        B, N = sun.shape[0], helio.shape[0]
        dev = action.device
        f32 = dict(dtype=torch.float32, device=dev)
        params = torch.empty(B, N, 4, **f32)
        actual = torch.empty(B, N, 3, **f32)
        refl = torch.empty(B * N, 3, **f32)
        ideal = bounds = angles = sums = None
        if want_aux:
            ideal = torch.empty(B, N, 3, **f32)
            bounds = torch.empty(B, N, **f32)
            angles = torch.empty(B, N, **f32)
            sums = torch.empty(2, **f32)
        with _Call("geom_fwd", dev):
            rc = lib.helio_geom_fwd(C.byref(scene), _ptr(helio), _ptr(sun), _ptr(action), _ptr(errs), B, N,
                                    _ptr(params), _ptr(actual), _ptr(refl), _ptr(ideal), _ptr(bounds), _ptr(angles),
                                    _ptr(sums), _ptr(workspace), workspace.numel() * workspace.element_size() if workspace is not None else 0,
                                    _stream())
        _lib.check(rc, "helio_geom_fwd")
        ctx.save_for_backward(action, sun, errs, helio)
        ctx.scene = scene
        ctx.set_materialize_grads(False)
        if want_aux:
            ctx.mark_non_differentiable(ideal)
        return params, actual, refl, ideal, bounds, angles, sums

    @staticmethod
    def backward(ctx, g_params, g_actual, g_refl, g_ideal, g_bounds, g_angles, g_sums):
        lib = _lib.load()
        action, sun, errs, helio = ctx.saved_tensors
        B, N = sun.shape[0], helio.shape[0]
        gs = [None if g is None else _cf(g) for g in (g_params, g_actual, g_refl, g_bounds, g_angles, g_sums)]
        g_action = torch.empty_like(action)
        with _Call("geom_bwd", action.device):
            rc = lib.helio_geom_bwd(C.byref(ctx.scene), _ptr(helio), _ptr(sun), _ptr(action), _ptr(errs), B, N,
                                    *[_ptr(g) for g in gs], _ptr(g_action), _stream())
        _lib.check(rc, "helio_geom_bwd")
        return g_action, None, None, None, None, None, None


class SplatFn(torch.autograd.Function):
    """K2/K3: params -> img[B,R,R]; backward returns the moments tensor in the slot of params.

    ``cull=True`` (shapes on the tcgen05 path only) contracts over the heliostats helio_cull keeps (see cull.cuh)."""

    @staticmethod
    def forward(ctx, params, R: int, width: float, height: float, impl: int, impl_bwd: int, cull: bool = False):
        lib = _lib.load()
        B, N = params.shape[0], params.shape[1]
        img = torch.empty(B, R, R, dtype=torch.float32, device=params.device)
        cull_ws = None
        if cull and impl != SPLAT_SIMT and impl_bwd != SPLAT_SIMT and int(lib.helio_step_partials_floats(B, N, R, impl)) > 0:
            cull_ws = _cull_workspace(lib, B, N, params.device)
        with _Call("splat_fwd", params.device):
            if cull_ws is not None:
                _lib.check(lib.helio_cull(_ptr(params), B, N, width, height, _ptr(cull_ws), cull_ws.numel() * 4, _stream()), "helio_cull")
                rc = lib.helio_splat_fwd_culled(_ptr(cull_ws), B, N, R, width, height, _ptr(img), _stream())
            else:
                rc = lib.helio_splat_fwd(_ptr(params), B, N, R, width, height, _ptr(img), impl, _stream())
        _lib.check(rc, "helio_splat_fwd")
        ctx.save_for_backward(params)
        ctx.cull_ws = cull_ws
        ctx.cfg = (R, width, height, impl_bwd)
        return img

    @staticmethod
    def backward(ctx, g_img):
        lib = _lib.load()
        (params,) = ctx.saved_tensors
        R, width, height, impl = ctx.cfg
        B, N = params.shape[0], params.shape[1]
        g_img = _cf(g_img)
        moments = torch.empty_like(params)
        with _Call("splat_bwd", params.device):
            if ctx.cull_ws is not None:
                rc = lib.helio_splat_bwd_culled(_ptr(ctx.cull_ws), _ptr(g_img), B, N, R, width, height, _ptr(moments), _stream())
            else:
                rc = lib.helio_splat_bwd(_ptr(params), _ptr(g_img), B, N, R, width, height, _ptr(moments), impl, _stream())
        _lib.check(rc, "helio_splat_bwd")
        return moments, None, None, None, None, None, None


def image_max(target: torch.Tensor) -> torch.Tensor:
    """tx[b] = max(target[b]).clamp_min(1e-6)  (test_environment.py:436)."""
    lib = _lib.load()
    target = _cf(target)
    B, R = target.shape[0], target.shape[-1]
    tx = torch.empty(B, dtype=torch.float32, device=target.device)
    with _Call("image_max", target.device):
        rc = lib.helio_image_max(_ptr(target), B, R, _ptr(tx), _stream())
    _lib.check(rc, "helio_image_max")
    return tx


def distance_maps(imgs: torch.Tensor, thr: float = 0.5) -> torch.Tensor:
    """Exact EDT of ``imgs > thr * max`` per image on the GPU (helio_distance_maps); replaces the host round trip
    through scipy.ndimage.distance_transform_edt of make_distance_maps (test_environment.py:92-97), bit for bit."""
    lib = _lib.load()
    imgs = _cf(imgs.detach())
    B, R = imgs.shape[0], imgs.shape[-1]
    assert imgs.dim() == 3 and imgs.shape[1] == R, "distance_maps expects [B, R, R]"
    out = torch.empty_like(imgs)
    nbytes = lib.helio_distance_maps_workspace_bytes(B, R)
    ws = torch.empty((nbytes + 3) // 4, dtype=torch.int32, device=imgs.device)
    with _Call("distance_maps", imgs.device):
        rc = lib.helio_distance_maps(_ptr(imgs), B, R, float(thr), _ptr(out), _ptr(ws), nbytes, _stream())
    _lib.check(rc, "helio_distance_maps")
    global _LAUNCHES
    _LAUNCHES += 2
    return out


class ImageLossFn(torch.autograd.Function):
    """K4: img -> per_img[B,3] = {sum diff^2, sum |diff| dmaps, sum |diff|}, diff=(img-target)/tx."""

    @staticmethod
    def forward(ctx, img, target, dmaps, tx):
        lib = _lib.load()
        img = _cf(img)
        B, R = img.shape[0], img.shape[-1]
        per_img = torch.empty(B, 3, dtype=torch.float32, device=img.device)
        with _Call("loss_fwd", img.device):
            rc = lib.helio_loss_fwd(_ptr(img), _ptr(target), _ptr(dmaps), _ptr(tx), B, R, _ptr(per_img), _stream())
        _lib.check(rc, "helio_loss_fwd")
        ctx.save_for_backward(img, target, dmaps, tx)
        return per_img

    @staticmethod
    def backward(ctx, g_per_img):
        lib = _lib.load()
        img, target, dmaps, tx = ctx.saved_tensors
        B, R = img.shape[0], img.shape[-1]
        g_per_img = _cf(g_per_img)
        g_img = torch.empty_like(img)
        with _Call("loss_bwd", img.device):
            rc = lib.helio_loss_bwd(_ptr(img), _ptr(target), _ptr(dmaps), _ptr(tx), _ptr(g_per_img), None, B, R,
                                    _ptr(g_img), _stream())
        _lib.check(rc, "helio_loss_bwd")
        return g_img, None, None, None


FUSE_LOSS_EPILOGUE = os.environ.get("HELIO_FUSE_LOSS", "0") == "1"


def _loss_partials(lib, B, N, R, impl, dev):
    """(scratch for the loss sums accumulated in the splat epilogue or None = separate loss_fwd pass, shape takes the
    tcgen05 splat).

    The loss fusion is off by default: measured on B200 the fused epilogue's per-row loads of target / dmaps add ~20 %
    to the L1 data-pipe wavefronts that bound the forward splat (+0.8 ms at N=2000, R=256, B=4096) and save only the
    0.48 ms loss pass.  The per-image maximum of the target (no extra loads) is always fused on the tcgen05 path."""
    n = int(lib.helio_step_partials_floats(B, N, R, impl))     # cheap host call; depends on the run-time pair mode: not cached
    if not FUSE_LOSS_EPILOGUE or n == 0:
        return None, n > 0
    return torch.empty(n, dtype=torch.float32, device=dev), True


_LAST_CULL = None


def _cull_workspace(lib, B, N, dev):
    """Compacted footprints + index map + per-sun counts of helio_cull (kept for the backward)."""
    global _LAST_CULL
    nbytes = int(lib.helio_cull_workspace_bytes(B, N))
    ws = torch.empty((nbytes + 3) // 4, dtype=torch.int32, device=dev)
    _LAST_CULL = (weakref.ref(ws), B, N)             # reporting only: must not keep the buffer alive
    return ws


def last_cull_kept_fraction():
    """Fraction of (sun, heliostat) pairs the most recent culled step kept (device sync; for reporting)."""
    if _LAST_CULL is None:
        return None
    ref, B, N = _LAST_CULL
    ws = ref()
    if ws is None:
        return None
    return float(ws[B * N * 5: B * N * 5 + B].sum().item()) / float(B * N)


def _step_fwd_kernels(render_target: bool, fused: bool, tc: bool = True) -> int:
    """Kernels helio_step_fwd enqueues: K1, K2, loss_pack (+ loss_fwd when the loss is not fused into K2's epilogue)
    and, with the target render, K1, K2 + either image_max (CUDA-core path) or the one-block fill of tx that seeds the
    maximum folded into the tcgen05 splat's epilogue."""
    return (3 if fused else 4) + (3 if render_target else 0)


class StepFn(torch.autograd.Function):
    """Whole HelioEnv.step forward / backward as ONE C-ABI call each (helio_step_fwd / helio_step_bwd).

    Same kernels and results as GeomFn -> SplatFn -> ImageLossFn composed by autograd; what it saves is
    host time (three Function nodes, ~25 small torch ops, 9 ctypes calls per step), which is all a small
    field costs (N=50, R=128, B=25: the kernels take a few microseconds).

    forward(action[B,N,3], sun, errs, helio, dmaps, scene, workspace, R, impl, impl_bwd, target, tx)
      -> img, packed[4] = {sum diff^2, sum |diff| dmaps, sum bounds, sum angles}, actual, refl, ideal,
         bounds, angles, per_img[B,3], target, tx
    ``target``/``tx`` None = render the target here (test_environment.py:429-436); else reuse them.
    """

    @staticmethod
    def forward(ctx, action, sun, errs, helio, dmaps, scene, workspace, R: int, impl: int, impl_bwd: int, target, tx, cull=False,
                want_com=False, img_out=None):
        """``want_com`` / ``img_out`` (encoder feed, helio_step_fwd_feed): the centre of mass of the noisy image and a second
        copy of it in ``img_out`` ([B,R,R] view, any batch stride, contiguous images) come out of the splat epilogue; the
        extra outputs com [B,2] (differentiable) and com_sums [B,3] are appended."""
        lib = _lib.load()
        B, N = sun.shape[0], helio.shape[0]
        dev = action.device
        f32 = dict(dtype=torch.float32, device=dev)
        render_target = target is None
        params = torch.empty(B, N, 4, **f32)
        actual, refl, ideal = torch.empty(B, N, 3, **f32), torch.empty(B * N, 3, **f32), torch.empty(B, N, 3, **f32)
        bounds, angles = torch.empty(B, N, **f32), torch.empty(B, N, **f32)
        img = torch.empty(B, R, R, **f32)
        per_img, packed = torch.empty(B, 3, **f32), torch.empty(4, **f32)
        scratch = None
        if render_target:
            target, tx = torch.empty(B, R, R, **f32), torch.empty(B, **f32)
            scratch = torch.empty(B * N * 10, **f32)     # params, actual, refl of the target render (discarded)
        partials, tc = _loss_partials(lib, B, N, R, impl, dev)
        cull_ws = _cull_workspace(lib, B, N, dev) if cull and tc else None
        feed = com = com_sums = feed_partials = None
        if want_com or img_out is not None:
            feed = Feed()
            feed.eps = 1e-12
            if want_com:
                com, com_sums = torch.empty(B, 2, **f32), torch.empty(B, 3, **f32)
                feed.com_coords, feed.com_sums = com.data_ptr(), com_sums.data_ptr()
            if img_out is not None:
                if tuple(img_out.shape) != (B, R, R) or img_out.dtype != torch.float32 or img_out.device != dev or \
                        img_out.stride(1) != R or img_out.stride(2) != 1:
                    raise ValueError("img_out must be a float32 [B,R,R] view on the env's device with contiguous images")
                feed.img2, feed.img2_batch_stride = img_out.data_ptr(), img_out.stride(0)
            nfl = int(lib.helio_step_partials_floats(B, N, R, impl))
            if nfl > 0:
                feed_partials = torch.empty(nfl, **f32)
                feed.partials, feed.partials_floats = feed_partials.data_ptr(), nfl
        with _Call("step_fwd", dev):
            rc = lib.helio_step_fwd_feed(
                C.byref(scene), _ptr(helio), _ptr(sun), _ptr(action), _ptr(errs), _ptr(dmaps), B, N, R, impl,
                1 if render_target else 0, _ptr(params), _ptr(actual), _ptr(refl), _ptr(ideal), _ptr(bounds), _ptr(angles),
                _ptr(img), _ptr(target), _ptr(tx), _ptr(per_img), _ptr(packed),
                _ptr(scratch), _ptr(scratch[4 * B * N:]) if render_target else None,
                _ptr(scratch[7 * B * N:]) if render_target else None, _ptr(partials), _ptr(cull_ws),
                _ptr(workspace), workspace.numel() * workspace.element_size(), None if feed is None else C.byref(feed), _stream())
        _lib.check(rc, "helio_step_fwd")
        global _LAUNCHES
        _LAUNCHES += _step_fwd_kernels(render_target, partials is not None and feed is None, tc) - 1 + (1 if cull_ws is not None else 0) \
            + (1 if want_com else 0)
        ctx.save_for_backward(action, sun, errs, helio, dmaps, params, img, target, tx, com_sums)
        ctx.cull_ws = cull_ws
        ctx.cfg = (scene, R, impl_bwd)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(ideal, target, tx)
        if want_com:
            ctx.mark_non_differentiable(com_sums)
        return img, packed, actual, refl, ideal, bounds, angles, per_img, target, tx, com, com_sums

    @staticmethod
    def backward(ctx, g_img_in, g_packed, g_actual, g_refl, g_ideal, g_bounds, g_angles, g_per_img, g_target, g_tx, g_com=None,
                 g_com_sums=None):
        lib = _lib.load()
        action, sun, errs, helio, dmaps, params, img, target, tx, com_sums = ctx.saved_tensors
        scene, R, impl = ctx.cfg
        B, N = sun.shape[0], helio.shape[0]
        global _LAUNCHES
        if g_com is not None:
            # adjoint of the centre of mass (helio_com_bwd): one elementwise pass producing dL/dimg, which joins whatever
            # arrives through obs['img'] and enters the loss backward / K3 as g_img_in
            g_img_com = torch.empty_like(img)
            _lib.check(lib.helio_com_bwd(_ptr(img), _ptr(com_sums), _ptr(_cf(g_com)), B, R, R, 1e-12, _ptr(g_img_com), _stream()),
                       "helio_com_bwd")
            _LAUNCHES += 1
            g_img_in = g_img_com if g_img_in is None else g_img_com.add_(g_img_in)
        gs = [None if g is None else _cf(g) for g in (g_packed, g_per_img, g_img_in, g_actual, g_refl, g_bounds, g_angles)]
        need_img = gs[0] is not None or gs[1] is not None
        need_splat = need_img or gs[2] is not None
        g_action = torch.empty_like(action)
        g_img = torch.empty_like(img) if need_img else None
        moments = torch.empty_like(params) if need_splat else None
        with _Call("step_bwd", action.device):
            rc = lib.helio_step_bwd(
                C.byref(scene), _ptr(helio), _ptr(sun), _ptr(action), _ptr(errs), _ptr(params), _ptr(img), _ptr(target),
                _ptr(dmaps), _ptr(tx), B, N, R, impl, *[_ptr(g) for g in gs], _ptr(ctx.cull_ws),
                _ptr(g_img), _ptr(moments), _ptr(g_action), _stream())
        _lib.check(rc, "helio_step_bwd")
        _LAUNCHES += (1 if need_img else 0) + (1 if need_splat else 0)
        return (g_action,) + (None,) * 14


def _host_slices(B: int, chunks: int, small_first: bool, quantum: int = 0):
    """[(b0, nb)] slices of the sun batch for overlapping host copies with kernels.  The one copy that cannot hide -- the first
    host->device slice of the forward, the last device->host slice of the backward -- is a quarter of the others.
    ``quantum`` (the SM count): slice sizes are rounded to whole waves of the persistent splat kernels (one sun = one tile of a
    CTA or CTA pair at R <= 256), so that only one slice carries a ragged last wave -- as the unsliced launch does -- instead
    of every slice (1260 suns on 74 CTA pairs = 17.03 waves: 18 are paid)."""
    chunks = max(1, min(int(chunks), B))
    if chunks == 1:
        return [(0, B)]
    small = max(1, int(round(0.25 * B / (chunks - 0.75))))
    q = int(quantum)
    if q > 1 and B >= 2 * q * chunks:
        small = max(q, small // q * q)
        body = max(q, int(round((B - small) / (chunks - 1) / q)) * q)
        while body > q and B - small - body * (chunks - 2) < q // 2:
            body -= q
        sizes = [body] * (chunks - 2) + [B - small - body * (chunks - 2)]      # the ragged remainder rides on one slice
    else:
        rest = B - small
        sizes = [rest // (chunks - 1) + (1 if i < rest % (chunks - 1) else 0) for i in range(chunks - 1)]
    sizes = [small] + sizes if small_first else sizes + [small]
    out, b0 = [], 0
    for nb in sizes:
        if nb > 0:
            out.append((b0, nb))
            b0 += nb
    return out


_SM_COUNT = {}


def _wave_quantum(dev) -> int:
    """SM count of the device (slice granularity of HostStepFn)."""
    idx = torch.device(dev).index
    idx = torch.cuda.current_device() if idx is None else idx
    if idx not in _SM_COUNT:
        _SM_COUNT[idx] = int(torch.cuda.get_device_properties(idx).multi_processor_count)
    return _SM_COUNT[idx]


class HostStepFn(torch.autograd.Function):
    """StepFn for an action that lives in HOST memory (the reference accepts host actions too: np.ndarray,
    test_environment.py:411-412), with the transfers overlapped instead of serialised:

      forward : the host->device copy of the action runs on a side stream while the main stream renders the target
                (which depends on the suns only); the noisy render starts when the copy lands.  With a cached target
                there is nothing to hide under, so the copy and the forward both run in ``chunks`` slices of the sun
                batch and slice k's kernels overlap slice k+1's copy;
      backward: runs in ``chunks`` slices of the sun batch; the device->host copy of a slice's action gradient
                overlaps the next slice's kernels.  Returns the gradient as a HOST tensor, pinned when the input was.

    The device copy of the action is returned as a differentiable output (obs['aux'] / monitor['normals'] are built from
    it, test_environment.py:424,:505); a gradient arriving through it is added to the action gradient before it leaves.

    Same kernels, same numbers as StepFn (the per-sun slices are independent; packed sums are formed over the whole
    batch in the forward)."""

    @staticmethod
    def forward(ctx, action_host, sun, errs, helio, dmaps, scene, workspace, R, impl, impl_bwd, target, tx, copy_stream,
                chunks: int, cull=False):
        lib = _lib.load()
        B, N = sun.shape[0], helio.shape[0]
        dev = sun.device
        f32 = dict(dtype=torch.float32, device=dev)
        main = torch.cuda.current_stream(dev)
        action = torch.empty(B, N, 3, **f32)
        params = torch.empty(B, N, 4, **f32)
        actual, refl, ideal_out = torch.empty(B, N, 3, **f32), torch.empty(B * N, 3, **f32), torch.empty(B, N, 3, **f32)
        bounds, angles = torch.empty(B, N, **f32), torch.empty(B, N, **f32)
        img = torch.empty(B, R, R, **f32)
        per_img, packed = torch.empty(B, 3, **f32), torch.empty(4, **f32)
        global _LAUNCHES
        partials, tc = _loss_partials(lib, B, N, R, impl, dev)
        cull_ws = _cull_workspace(lib, B, N, dev) if cull and tc else None
        chunks = max(1, min(int(chunks), B))
        if cull_ws is not None or partials is not None:
            chunks = 1                                   # the culled lists / fused-loss partials are indexed by the whole batch
        render_target = target is None
        # With the target to render the whole copy hides under it; with a cached target the forward itself runs in
        # slices of the sun batch, each starting when its slice of the action has landed.
        fwd_slices = _host_slices(B, 1 if render_target else chunks, small_first=True, quantum=_wave_quantum(dev))
        fwd_chunks = len(fwd_slices)
        landed = []
        action_host3 = action_host.reshape(B, N, 3)
        copy_stream.wait_stream(main)                    # `action` is allocated on main's pool: order its first use
        with torch.cuda.stream(copy_stream):
            for b0, nb in fwd_slices:
                action.narrow(0, b0, nb).copy_(action_host3.narrow(0, b0, nb), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                landed.append(ev)
        nws = workspace.numel() * workspace.element_size()
        sl = lambda t, b0, nb: None if t is None else t.narrow(0, b0, nb)
        with _Call("step_fwd_host", dev):
            if render_target:                            # phase 1, target only: it does not need the action
                target, tx = torch.empty(B, R, R, **f32), torch.empty(B, **f32)
                scratch = torch.empty(B * N * 10, **f32)
                rc = lib.helio_step_fwd(
                    C.byref(scene), _ptr(helio), _ptr(sun), None, None, None, B, N, R, impl, 1, None, None, None, None, None, None,
                    None, _ptr(target), _ptr(tx), None, None, _ptr(scratch), _ptr(scratch[4 * B * N:]), _ptr(scratch[7 * B * N:]),
                    _ptr(partials), None, _ptr(workspace), nws, _stream())
                _lib.check(rc, "helio_step_fwd (target)")
            if fwd_chunks == 1:
                main.wait_event(landed[0])
                rc = lib.helio_step_fwd(
                    C.byref(scene), _ptr(helio), _ptr(sun), _ptr(action), _ptr(errs), _ptr(dmaps), B, N, R, impl, 0,
                    _ptr(params), _ptr(actual), _ptr(refl), _ptr(ideal_out), _ptr(bounds), _ptr(angles), _ptr(img), _ptr(target),
                    _ptr(tx), _ptr(per_img), _ptr(packed), None, None, None, _ptr(partials), _ptr(cull_ws), _ptr(workspace), nws,
                    _stream())
                _lib.check(rc, "helio_step_fwd")
            else:
                # per-slice forward; the per-image sums are the same numbers as in one call and are packed over the whole
                # batch afterwards (bit-identical mse / dist); the bound / alignment sums are added slice by slice
                parts = torch.empty(len(landed), 4, **f32)
                refl3 = refl.view(B, N, 3)
                for k, (b0, nb) in enumerate(fwd_slices):
                    main.wait_event(landed[k])
                    rc = lib.helio_step_fwd(
                        C.byref(scene), _ptr(helio), _ptr(sl(sun, b0, nb)), _ptr(sl(action, b0, nb)), _ptr(sl(errs, b0, nb)),
                        _ptr(sl(dmaps, b0, nb)), nb, N, R, impl, 0, _ptr(sl(params, b0, nb)), _ptr(sl(actual, b0, nb)),
                        _ptr(sl(refl3, b0, nb)), _ptr(sl(ideal_out, b0, nb)), _ptr(sl(bounds, b0, nb)), _ptr(sl(angles, b0, nb)),
                        _ptr(sl(img, b0, nb)), _ptr(sl(target, b0, nb)), _ptr(sl(tx, b0, nb)), _ptr(sl(per_img, b0, nb)), _ptr(parts[k]),
                        None, None, None, None, None, _ptr(workspace), nws, _stream())
                    _lib.check(rc, "helio_step_fwd (slice)")
                    _LAUNCHES += _step_fwd_kernels(False, False, tc) if k else 0
                _lib.check(lib.helio_loss_pack(_ptr(per_img), B, _ptr(packed), _stream()), "helio_loss_pack")
                packed[2:].copy_(parts[:, 2:].sum(0))
                _LAUNCHES += 1
        _LAUNCHES += _step_fwd_kernels(render_target, partials is not None, tc) - 1 + (1 if cull_ws is not None else 0)
        ctx.cull_ws = cull_ws
        ctx.save_for_backward(action, sun, errs, helio, dmaps, params, img, target, tx)
        ctx.cfg = (scene, R, impl_bwd, copy_stream, chunks, action_host.shape, action_host.is_pinned())
        ctx.set_materialize_grads(False)
        action_dev = action.view(B, N, 3)             # handed back so that obs['aux'] / monitor['normals'] live on the device
        ctx.mark_non_differentiable(ideal_out, target, tx)
        return img, packed, actual, refl, ideal_out, bounds, angles, per_img, target, tx, action_dev

    @staticmethod
    def backward(ctx, g_img_in, g_packed, g_actual, g_refl, g_ideal, g_bounds, g_angles, g_per_img, g_target, g_tx, g_adev):
        lib = _lib.load()
        action, sun, errs, helio, dmaps, params, img, target, tx = ctx.saved_tensors
        scene, R, impl, copy_stream, chunks, host_shape, pinned = ctx.cfg
        B, N = sun.shape[0], helio.shape[0]
        dev = action.device
        gs = [None if g is None else _cf(g) for g in (g_packed, g_per_img, g_img_in, g_actual, g_refl, g_bounds, g_angles)]
        need_img = gs[0] is not None or gs[1] is not None
        need_splat = need_img or gs[2] is not None
        g_action = torch.empty_like(action)
        g_img = torch.empty_like(img) if need_img else None
        moments = torch.empty_like(params) if need_splat else None
        h_grad = torch.empty(B, N, 3, dtype=torch.float32, device="cpu", pin_memory=pinned)   # explicit: callers may set a CUDA default device
        main = torch.cuda.current_stream(dev)
        sl = lambda t, b0, nb: None if t is None else t.narrow(0, b0, nb)
        global _LAUNCHES
        with _Call("step_bwd_host", dev):
            for b0, nb in _host_slices(B, chunks, small_first=False, quantum=_wave_quantum(g_action.device)):
                refl_g = None if gs[4] is None else gs[4].view(B, N, 3).narrow(0, b0, nb)
                rc = lib.helio_step_bwd(
                    C.byref(scene), _ptr(helio), _ptr(sl(sun, b0, nb)), _ptr(sl(action, b0, nb)), _ptr(sl(errs, b0, nb)),
                    _ptr(sl(params, b0, nb)), _ptr(sl(img, b0, nb)), _ptr(sl(target, b0, nb)), _ptr(sl(dmaps, b0, nb)),
                    _ptr(sl(tx, b0, nb)), nb, N, R, impl, _ptr(gs[0]), _ptr(sl(gs[1], b0, nb)), _ptr(sl(gs[2], b0, nb)),
                    _ptr(sl(gs[3], b0, nb)), _ptr(refl_g), _ptr(sl(gs[5], b0, nb)), _ptr(sl(gs[6], b0, nb)), _ptr(ctx.cull_ws),
                    _ptr(sl(g_img, b0, nb)), _ptr(sl(moments, b0, nb)), _ptr(sl(g_action, b0, nb)), _stream())
                _lib.check(rc, "helio_step_bwd")
                _LAUNCHES += (1 if need_img else 0) + (1 if need_splat else 0) + (1 if b0 else 0)
                if g_adev is not None:                # gradient that arrived through the device copy (aux / monitor['normals'])
                    g_action.narrow(0, b0, nb).add_(g_adev.reshape(B, N, 3).narrow(0, b0, nb))
                done = torch.cuda.Event()
                done.record(main)
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(done)
                    h_grad.narrow(0, b0, nb).copy_(g_action.narrow(0, b0, nb), non_blocking=True)
        main.wait_stream(copy_stream)                 # g_action is freed on main: keep the allocator's ordering
        copy_stream.synchronize()                     # the caller owns a host tensor: it must be complete
        return (h_grad.view(host_shape),) + (None,) * 14
