"""ctypes binding of libhelio_sm100.so (the C ABI declared in include/helio_b200.h).

There is no CPU fallback: if the shared library is missing or the device is not a cc-10.x GPU the
product path raises.  ``build()`` compiles the library in-tree with nvcc (sm_100a only).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HELIO_LIB_PATH") or os.path.join(_HERE, "libhelio_sm100.so")   # override: experiments only
CSRC = os.path.join(_HERE, "csrc")
ABI_VERSION = 5

SPLAT_AUTO, SPLAT_SIMT, SPLAT_TC = 0, 1, 2
BWD_PREC_DEFAULT = 1     # helio_set_bwd_precision: 1 = f16x3 (K = 64 per stage) inside helio_step_bwd, 0 = 3xTF32

EXPORTS = (
    "helio_abi_version", "helio_last_error", "helio_device_ok", "helio_set_tc_pair_mode", "helio_set_fwd_precision", "helio_set_bwd_precision", "helio_geom_workspace_bytes", "helio_geom_fwd",
    "helio_geom_bwd", "helio_splat_fwd", "helio_splat_bwd", "helio_image_max", "helio_loss_fwd", "helio_loss_bwd",
    "helio_profile_enable", "helio_profile_count", "helio_profile_get",
    "helio_distance_maps_workspace_bytes", "helio_distance_maps",
    "helio_com_fwd", "helio_com_bwd",
    "helio_cull_workspace_bytes", "helio_cull", "helio_splat_fwd_culled", "helio_splat_bwd_culled",
    "helio_loss_bwd_packed", "helio_loss_pack", "helio_step_partials_floats", "helio_step_fwd", "helio_step_bwd",
    "helio_splat_fwd_feed", "helio_step_fwd_feed", "helio_tc_clock_mhz",
)


class Scene(C.Structure):
    """helio_scene_t (include/helio_b200.h)."""
    _fields_ = [
        ("target_pos", C.c_float * 3), ("target_normal", C.c_float * 3), ("plane_u", C.c_float * 3),
        ("plane_v", C.c_float * 3), ("width", C.c_float), ("height", C.c_float), ("sigma_scale", C.c_float),
        ("bnd_targ_pos", C.c_float * 3), ("bnd_targ_norm", C.c_float * 3), ("bnd_u", C.c_float * 3),
        ("bnd_v", C.c_float * 3), ("bnd_width", C.c_float), ("bnd_height", C.c_float),
    ]


class Feed(C.Structure):
    """helio_feed_t (include/helio_b200.h): encoder-feed outputs of the splat epilogue."""
    _fields_ = [("img2", C.c_void_p), ("img2_batch_stride", C.c_int64), ("eps", C.c_float), ("com_coords", C.c_void_p),
                ("com_sums", C.c_void_p), ("partials", C.c_void_p), ("partials_floats", C.c_int64)]


class HelioLibError(RuntimeError):
    pass


_lock = threading.Lock()
_lib = None


def build(verbose: bool = False) -> str:
    """Compile csrc/ -> libhelio_sm100.so for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC], capture_output=True, text=True)
    if r.returncode != 0:
        raise HelioLibError("building libhelio_sm100.so failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)
    return LIB_PATH


def _declare(lib):
    p, i, f, i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64
    sp = C.POINTER(Scene)
    lib.helio_abi_version.restype = i
    lib.helio_abi_version.argtypes = []
    lib.helio_last_error.restype = C.c_char_p
    lib.helio_last_error.argtypes = []
    lib.helio_device_ok.restype = i
    lib.helio_device_ok.argtypes = []
    lib.helio_set_tc_pair_mode.restype = i
    lib.helio_set_tc_pair_mode.argtypes = [i]
    lib.helio_set_fwd_precision.restype = i
    lib.helio_set_fwd_precision.argtypes = [i]
    lib.helio_set_bwd_precision.restype = i
    lib.helio_set_bwd_precision.argtypes = [i]
    lib.helio_profile_enable.restype = i
    lib.helio_profile_enable.argtypes = [i]
    lib.helio_profile_count.restype = i
    lib.helio_profile_count.argtypes = []
    lib.helio_profile_get.restype = i
    lib.helio_profile_get.argtypes = [i, C.POINTER(C.c_char_p), C.POINTER(C.c_float)]
    lib.helio_distance_maps_workspace_bytes.restype = i64
    lib.helio_distance_maps_workspace_bytes.argtypes = [i, i]
    lib.helio_distance_maps.restype = i
    lib.helio_distance_maps.argtypes = [p, i, i, f, p, p, i64, p]
    lib.helio_com_fwd.restype = i
    lib.helio_com_fwd.argtypes = [p, i, i, i, f, p, p, p]
    lib.helio_com_bwd.restype = i
    lib.helio_com_bwd.argtypes = [p, p, p, i, i, i, f, p, p]
    lib.helio_cull_workspace_bytes.restype = i64
    lib.helio_cull_workspace_bytes.argtypes = [i, i]
    lib.helio_cull.restype = i
    lib.helio_cull.argtypes = [p, i, i, f, f, p, i64, p]
    lib.helio_splat_fwd_culled.restype = i
    lib.helio_splat_fwd_culled.argtypes = [p, i, i, i, f, f, p, p]
    lib.helio_splat_bwd_culled.restype = i
    lib.helio_splat_bwd_culled.argtypes = [p, p, i, i, i, f, f, p, p]
    lib.helio_geom_workspace_bytes.restype = i64
    lib.helio_geom_workspace_bytes.argtypes = [i, i]
    lib.helio_geom_fwd.restype = i
    lib.helio_geom_fwd.argtypes = [sp, p, p, p, p, i, i, p, p, p, p, p, p, p, p, i64, p]
    lib.helio_geom_bwd.restype = i
    lib.helio_geom_bwd.argtypes = [sp, p, p, p, p, i, i, p, p, p, p, p, p, p, p]
    lib.helio_splat_fwd.restype = i
    lib.helio_splat_fwd.argtypes = [p, i, i, i, f, f, p, i, p]
    lib.helio_splat_bwd.restype = i
    lib.helio_splat_bwd.argtypes = [p, p, i, i, i, f, f, p, i, p]
    lib.helio_image_max.restype = i
    lib.helio_image_max.argtypes = [p, i, i, p, p]
    lib.helio_loss_fwd.restype = i
    lib.helio_loss_fwd.argtypes = [p, p, p, p, i, i, p, p]
    lib.helio_loss_bwd.restype = i
    lib.helio_loss_bwd.argtypes = [p, p, p, p, p, p, i, i, p, p]
    lib.helio_loss_bwd_packed.restype = i
    lib.helio_loss_bwd_packed.argtypes = [p, p, p, p, p, p, p, i, i, p, p]
    lib.helio_loss_pack.restype = i
    lib.helio_loss_pack.argtypes = [p, i, p, p]
    lib.helio_step_fwd.restype = i
    lib.helio_step_fwd.argtypes = [sp, p, p, p, p, p, i, i, i, i, i] + [p] * 16 + [p, i64, p]
    lib.helio_step_fwd_feed.restype = i
    lib.helio_step_fwd_feed.argtypes = [sp, p, p, p, p, p, i, i, i, i, i] + [p] * 16 + [p, i64, C.POINTER(Feed), p]
    lib.helio_splat_fwd_feed.restype = i
    lib.helio_splat_fwd_feed.argtypes = [p, i, i, i, f, f, p, i, C.POINTER(Feed), p]
    lib.helio_tc_clock_mhz.restype = i
    lib.helio_tc_clock_mhz.argtypes = [i, C.POINTER(C.c_float)]
    lib.helio_step_partials_floats.restype = i64
    lib.helio_step_partials_floats.argtypes = [i, i, i, i]
    lib.helio_step_bwd.restype = i
    lib.helio_step_bwd.argtypes = [sp] + [p] * 9 + [i, i, i, i] + [p] * 8 + [p, p, p, p]


def load(build_if_missing: bool = False):
    """Return the loaded library; raise HelioLibError if it cannot be loaded (no fallback)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            if not build_if_missing:
                raise HelioLibError(
                    f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "or `make -C doodle_b200/csrc`.  doodle_b200 has no CPU / PyTorch fallback.")
            build()
        try:
            lib = C.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover
            raise HelioLibError(f"cannot load {LIB_PATH}: {e}") from e
        for name in EXPORTS:
            if not hasattr(lib, name):
                raise HelioLibError(f"{LIB_PATH} does not export {name}")
        _declare(lib)
        if lib.helio_abi_version() != ABI_VERSION:
            raise HelioLibError(f"ABI mismatch: library {lib.helio_abi_version()} != binding {ABI_VERSION}")
        _lib = lib
        return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().helio_last_error().decode(errors="replace")
        raise HelioLibError(f"{what} failed (rc={rc}): {msg}")
