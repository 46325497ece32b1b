"""CPU-side checks: the C-ABI library loads and exports every declared symbol, the product refuses to
run without a GPU, and the host-side helpers of the env shim match the reference (golden fixtures)."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden


def test_library_exports_every_declared_symbol():
    from doodle_b200 import _lib
    lib = _lib.load(build_if_missing=True)
    header = open(os.path.join(ROOT, "include", "helio_b200.h")).read()
    declared = set(re.findall(r"HELIO_API[^;(]*?\b(helio_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.helio_abi_version() == _lib.ABI_VERSION
    assert int(re.search(r"#define HELIO_ABI_VERSION (\d+)", header).group(1)) == _lib.ABI_VERSION


def test_scene_struct_layout_matches_header():
    import ctypes as C
    from doodle_b200._lib import Scene
    assert C.sizeof(Scene) == 29 * 4          # 4*3 + 3 + 4*3 + 2 floats, no padding
    assert Scene.sigma_scale.offset == 14 * 4 and Scene.bnd_width.offset == 27 * 4


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_no_cpu_fallback():
    from doodle_b200 import HelioEnv, HelioField
    h = torch.rand(3, 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        HelioField(h, torch.zeros(3), (1., 1.), torch.tensor([0., 1., 0.]), device="cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        HelioEnv(h, torch.zeros(3), (1., 1.), torch.tensor([0., 1., 0.]), device="cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        HelioEnv(h, torch.zeros(3), (1., 1.), torch.tensor([0., 1., 0.]), device="cuda")   # CUDA unavailable here


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "doodle_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("no CPU", ""), fn


def test_host_helpers_match_reference():
    from doodle_b200 import azimuth_elevation_to_primary_direction, make_distance_maps, sample_cone_directions
    g = load_golden("host_helpers")
    torch.manual_seed(7)
    axis = azimuth_elevation_to_primary_direction(45.0, 45.0)
    np.testing.assert_allclose(axis.numpy(), g["axis"], rtol=1e-6)
    dirs = sample_cone_directions(9, axis, 2.0, force_upper_hemisphere=True)
    np.testing.assert_allclose(dirs.numpy(), g["dirs"], rtol=1e-5, atol=1e-6)
    axis2 = azimuth_elevation_to_primary_direction(10.0, 89.9)
    dirs2 = sample_cone_directions(5, axis2, 2.0, force_upper_hemisphere=True)   # near-vertical axis: helper switch
    np.testing.assert_allclose(dirs2.numpy(), g["dirs2"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(make_distance_maps(torch.as_tensor(g["imgs"])).numpy(), g["dmaps"], rtol=1e-6)
    cosang = (dirs * axis).sum(1)
    assert float(cosang.min()) >= np.cos(np.radians(2.0)) - 1e-6


def test_api_signatures_match_reference():
    """Constructor / method signatures kept verbatim."""
    import inspect
    from doodle_b200 import HelioEnv, HelioField
    p = list(inspect.signature(HelioField.__init__).parameters)
    assert p[:11] == ["self", "heliostat_positions", "target_position", "target_area", "target_normal", "error_scale_mrad",
                 "sigma_scale", "initial_action_noise", "resolution", "device", "max_batch_size"]
    d = {k: v.default for k, v in inspect.signature(HelioField.__init__).parameters.items()}
    assert (d["error_scale_mrad"], d["sigma_scale"], d["initial_action_noise"], d["resolution"], d["max_batch_size"]) == (1.0, 0.01, 0.01, 100, 25)
    assert list(inspect.signature(HelioField.render).parameters) == ["self", "sun_position", "action", "ideal_normals", "show_spillage", "monitor"]
    p = list(inspect.signature(HelioEnv.__init__).parameters)
    assert p[:19] == ["self", "heliostat_pos", "targ_pos", "targ_area", "targ_norm", "sigma_scale", "error_scale_mrad",
                      "initial_action_noise", "resolution", "batch_size", "device", "new_sun_pos_every_reset",
                      "new_errors_every_reset", "use_error_mask", "error_mask_ratio", "exponential_risk", "single_sun",
                      "azimuth", "elevation"]
    d = {k: v.default for k, v in inspect.signature(HelioEnv.__init__).parameters.items()}
    assert (d["sigma_scale"], d["error_scale_mrad"], d["resolution"], d["batch_size"], d["device"]) == (0.1, 180.0, 128, 25, "cuda")
    for m in ("reset", "step", "set_sun_pos", "seed"):
        assert callable(getattr(HelioEnv, m))
    for m in ("reset_errors", "_sample_error_angles", "calculate_ideal_normals", "init_actions", "render"):
        assert callable(getattr(HelioField, m))


def test_dropin_modules_resolve():
    import importlib
    import sys
    sys.path.insert(0, os.path.join(ROOT, "dropin"))
    try:
        a = importlib.import_module("newenv_rl_test_multi_error")
        from doodle_b200 import HelioField
        assert a.HelioField is HelioField
    finally:
        sys.path.remove(os.path.join(ROOT, "dropin"))
        sys.modules.pop("newenv_rl_test_multi_error", None)


def test_header_is_plain_c():
    """include/helio_b200.h is the C ABI: it must compile as C99 on its own (no C++ constructs, no CUDA headers)."""
    import shutil, subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    r = subprocess.run([gcc, "-fsyntax-only", "-x", "c", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror",
                        os.path.join(ROOT, "include", "helio_b200.h")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_c_program_can_bind_the_library(tmp_path):
    """A plain C host (what a non-Python maintainer would write) dlopens libhelio_sm100.so, resolves every entry point
    declared in the header and calls the ones that need no GPU."""
    import re, shutil, subprocess
    from doodle_b200 import _lib
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    _lib.load(build_if_missing=True)
    header = open(os.path.join(ROOT, "include", "helio_b200.h")).read()
    names = sorted(set(re.findall(r"HELIO_API[^;(]*?\b(helio_[a-z0-9_]+)\s*\(", header)))
    src = tmp_path / "bind.c"
    src.write_text('''
#include <dlfcn.h>
#include <stdio.h>
#include "helio_b200.h"
int main(int argc, char** argv) {
    void* h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
    if (!h) { fprintf(stderr, "dlopen: %s\\n", dlerror()); return 2; }
    const char* names[] = {''' + ", ".join(f'"{n}"' for n in names) + '''};
    for (unsigned i = 0; i < sizeof names / sizeof names[0]; ++i)
        if (!dlsym(h, names[i])) { fprintf(stderr, "missing %s\\n", names[i]); return 3; }
    int (*ver)(void) = (int (*)(void))dlsym(h, "helio_abi_version");
    long long (*wsb)(int, int) = (long long (*)(int, int))dlsym(h, "helio_geom_workspace_bytes");
    if (ver() != HELIO_ABI_VERSION) return 4;
    if (wsb(4, 100) <= 0 || wsb(0, 100) != 0) return 5;
    printf("abi %d, %u symbols\\n", ver(), (unsigned)(sizeof names / sizeof names[0]));
    return 0;
}
''')
    exe = tmp_path / "bind"
    r = subprocess.run([gcc, "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-ldl"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe), _lib.LIB_PATH], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert f"abi {_lib.ABI_VERSION}" in r.stdout


@pytest.mark.skipif(not os.path.exists("/root/reference/train_with_env.py"), reason="needs a checkout of the reference (build container only)")
@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check: on a GPU the trainer would start training")
def test_reference_trainer_runs_unmodified_up_to_our_env():
    """scripts/run_reference_trainer.sh: the reference's own train_with_env.py, unmodified, must import through the
    stand-ins and the dropin/ shims and reach OUR HelioEnv (which refuses to run without a GPU) -- i.e. the module
    shadowing really takes effect although python puts the script's directory first on sys.path."""
    import subprocess
    r = subprocess.run(["bash", os.path.join(ROOT, "scripts", "run_reference_trainer.sh"), "/root/reference"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode != 0
    assert "doodle_b200/env.py" in r.stderr and "no CPU fallback" in r.stderr, r.stderr[-2000:]
    assert "/root/reference/test_environment.py" not in r.stderr          # the reference's env was NOT the one imported


def test_angular_action_space_matches_reference_rotation():
    """angles_to_normals vs the reference's rotate_normals_batch applied to north-pointing normals
    (newenv/test_environment_angular.py:205-214), values and autograd (tests/golden/angular.npz)."""
    from doodle_b200 import angles_to_normals
    g = load_golden("angular")
    a = torch.as_tensor(g["angles"]).requires_grad_(True)
    n = angles_to_normals(a, g["normals"].shape[1])
    np.testing.assert_allclose(n.detach().numpy(), g["normals"], rtol=1e-6, atol=1e-7)
    gr, = torch.autograd.grad((n * torch.as_tensor(g["w"])).sum(), a)
    np.testing.assert_allclose(gr.numpy(), g["grad"], rtol=1e-5, atol=1e-9)
    n3 = angles_to_normals(torch.as_tensor(g["angles"]).view(3, -1, 2), g["normals"].shape[1])    # [B,N,2] layout
    assert torch.equal(n3, n.detach())


def test_host_slices_cover_the_batch_in_whole_waves():
    """HostStepFn's slices of the sun batch: a partition of [0, B), the exposed slice (first of the forward, last of the backward)
    a quarter of the others, and -- for batches of many waves -- every slice but one a multiple of the SM count."""
    from doodle_b200.functional import _host_slices
    for B in (1, 5, 25, 512, 1184, 2368, 4096, 16384):
        for chunks in (1, 3, 4, 8):
            for small_first in (True, False):
                for q in (0, 148):
                    sl = _host_slices(B, chunks, small_first, q)
                    assert sl[0][0] == 0 and all(nb > 0 for _, nb in sl)
                    assert all(sl[i][0] + sl[i][1] == sl[i + 1][0] for i in range(len(sl) - 1))
                    assert sl[-1][0] + sl[-1][1] == B and len(sl) <= chunks
                    if len(sl) > 1:
                        small = sl[0][1] if small_first else sl[-1][1]
                        assert small <= min(nb for _, nb in sl)
                    if q and B >= 2 * q * chunks and chunks > 1:
                        assert sum(1 for _, nb in sl if nb % q) <= 1, (B, chunks, sl)
