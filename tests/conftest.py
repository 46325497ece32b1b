import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    if "w_seed" in g and "w_img" not in g:       # large fixtures store the seed of the image cotangent (oracle/make_golden.py)
        g["w_img"] = np.random.RandomState(int(g["w_seed"])).standard_normal(g["img"].shape).astype(np.float32)
    return g


@pytest.fixture
def golden():
    return load_golden


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


# ---- gradient parity (BASELINE.json north_star: "action gradients within 1e-3 relative") ------------------------------
# Elementwise, not a max-norm: every component must satisfy
#     |g - g_ref| <= 1e-3 |g_ref| + 1e-3 median|g_ref| + 1e-5 max_c |g_ref[b,n,c]|
# so a component 100x smaller than the largest cannot hide a 10 % error.  The median (taken over the components that are
# not numerically zero: with large orientation errors half of the heliostats miss the receiver and their gradient is 0 or
# denormal) is the absolute floor.  The last term applies to [.., 3] gradients: the three components of one heliostat's
# gradient leave ONE adjoint chain, and a component that is a cancellation residue of its siblings (e.g. [2719, -2362, 3.5])
# carries the fp32 roundoff of the whole vector -- the reference's own fp32 autograd differs from fp64 there by ~5e-6 of the
# vector norm.  1e-5 of a heliostat's own vector is 100x tighter than the old 1e-3 of the global maximum.
GRAD_RTOL = 1e-3
FP64_TIEBREAKS = {"n": 0, "cases": []}


def grad_excess(g, ref, rtol=GRAD_RTOL, extra_rtol=None):
    """max over components of |g - ref| / ((rtol + extra) |ref| + rtol median|ref| + 1e-5 |ref[b,n,:]|_inf); <= 1 passes.
    ``extra_rtol`` (broadcastable to ref's leading dims, e.g. [B,N] for a [B,N,3] gradient): conditioning-aware slack."""
    g = np.asarray(g, np.float64)
    ref = np.asarray(ref, np.float64).reshape(g.shape)
    mx = float(np.abs(ref).max()) if ref.size else 0.0
    live = np.abs(ref)[np.abs(ref) > 1e-6 * mx]
    med = float(np.median(live)) if live.size else mx
    r = rtol
    if extra_rtol is not None:
        e = np.asarray(extra_rtol, np.float64)
        r = rtol + e.reshape(e.shape + (1,) * (ref.ndim - e.ndim))
    tol = r * np.abs(ref) + rtol * med
    if ref.ndim >= 2 and ref.shape[-1] == 3:
        tol = tol + 1e-5 * np.abs(ref).max(axis=-1, keepdims=True)
    return float((np.abs(g - ref) / np.maximum(tol, 1e-300)).max()) if g.size else 0.0


def acos_grad_slack(angles_mrad):
    """Relative slack for d(alignment)/d(action) per (sun, heliostat): angle = 1000 acos(dot) has derivative
    -1000 / sqrt(1 - dot^2); a 2-ulp fp32 rounding of dot (2.4e-7 near 1) changes it by 2.4e-7 / angle^2 relative, which is
    percent-level for mirrors aligned to a few mrad -- in the reference's own fp32 autograd as much as in any other fp32
    evaluation.  Floor: the acos clamp at 0.3453 mrad (test_environment.py:132-155)."""
    th = np.maximum(np.asarray(angles_mrad, np.float64) * 1e-3, 3.4e-4)
    return 2.4e-7 / (th * th)


def assert_grad_close(g, ref, ref64=None, what="", rtol=GRAD_RTOL, extra_rtol=None):
    """Elementwise gradient check against the fp32 reference; ``ref64`` (fp64 oracle of the same inputs) decides when two
    fp32 results disagree near tolerance.  Tie-break uses are counted and reported at the end of the session."""
    ex = grad_excess(g, ref, rtol, extra_rtol)
    if ex <= 1.0:
        return
    if ref64 is not None:
        ex64 = grad_excess(g, ref64, rtol, extra_rtol)
        if ex64 <= 1.0:
            FP64_TIEBREAKS["n"] += 1
            FP64_TIEBREAKS["cases"].append(f"{what}: excess vs fp32 reference {ex:.2f}, vs fp64 oracle {ex64:.2f}")
            return
        raise AssertionError(f"gradient {what}: elementwise excess {ex:.3f} vs fp32 reference, {ex64:.3f} vs fp64 oracle (must be <= 1)")
    raise AssertionError(f"gradient {what}: elementwise excess {ex:.3f} (must be <= 1; |g-ref| <= {rtol}|ref| + {rtol} median|ref| + 1e-5 |ref[b,n,:]|inf)")


def pytest_terminal_summary(terminalreporter):
    n = FP64_TIEBREAKS["n"]
    terminalreporter.write_line(f"fp64 tie-breaks used by gradient checks: {n}")
    for c in FP64_TIEBREAKS["cases"]:
        terminalreporter.write_line("  " + c)
