"""Probe of the tcgen05 splat kernels against the SIMT path (run under gpurun with a timeout).

    python scripts/gpu_tc_probe.py [--quick] [--no-time]
"""
import argparse
import ctypes as C
import sys

sys.path.insert(0, ".")
import torch

from doodle_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
ap.add_argument("--no-time", action="store_true")
ap.add_argument("--B", type=int, default=4096)
a = ap.parse_args()

lib = _lib.load()
dev = torch.device("cuda:0")
P = lambda t: C.c_void_p(t.data_ptr())


def make_params(B, N, seed=1):
    g = torch.Generator(device=dev).manual_seed(seed)
    p = torch.empty(B, N, 4, device=dev)
    p[..., 0] = (torch.rand(B, N, device=dev, generator=g) - 0.5) * 10
    p[..., 1] = (torch.rand(B, N, device=dev, generator=g) - 0.5) * 10
    p[..., 2] = 1.4427 / (2 * (0.8 + 0.6 * torch.rand(B, N, device=dev, generator=g)) ** 2)
    p[..., 3] = 0.9 + 0.1 * torch.rand(B, N, device=dev, generator=g)
    return p


def fwd(p, R, impl):
    B, N = p.shape[:2]
    img = torch.full((B, R, R), float("nan"), device=dev)
    rc = lib.helio_splat_fwd(P(p), B, N, R, 15.0, 15.0, P(img), impl, None)
    assert rc == 0, lib.helio_last_error()
    torch.cuda.synchronize()
    return img


def bwd(p, g, R, impl):
    B, N = p.shape[:2]
    mom = torch.full((B, N, 4), float("nan"), device=dev)
    rc = lib.helio_splat_bwd(P(p), P(g), B, N, R, 15.0, 15.0, P(mom), impl, None)
    assert rc == 0, lib.helio_last_error()
    torch.cuda.synchronize()
    return mom


shapes = [(1, 32, 256), (2, 40, 256), (3, 300, 256), (2, 64, 128), (5, 100, 200), (2, 130, 64), (3, 37, 100), (1, 5, 33),
          (2, 70, 512), (300, 64, 256)]
if a.quick:
    shapes = shapes[:4]
ok = True
for (B, N, R) in shapes:
    p = make_params(B, N)
    s, t = fwd(p, R, 1), fwd(p, R, 2)
    ferr = ((s - t).abs() / (1e-6 + 1e-4 * s.abs())).max().item()
    g = torch.randn(B, R, R, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
    ms, mt = bwd(p, g, R, 1), bwd(p, g, R, 2)
    scale = ms.abs().amax(dim=(0, 1))
    berr = ((ms - mt).abs().amax(dim=(0, 1)) / scale).max().item()
    nan = int(torch.isnan(t).sum().item() + torch.isnan(mt).sum().item())
    flag = "ok" if (ferr < 1.0 and berr < 1e-4 and nan == 0) else "FAIL"
    ok &= flag == "ok"
    print(f"B={B} N={N} R={R}: fwd tol-ratio {ferr:.3f}  bwd max-rel {berr:.2e}  nan {nan}  {flag}", flush=True)
print("PARITY", "OK" if ok else "FAILED", flush=True)

if not a.no_time:
    B, N, R = a.B, 2000, 256
    p = make_params(B, N)
    img = torch.empty(B, R, R, device=dev)
    g = torch.randn(B, R, R, device=dev)
    mom = torch.empty(B, N, 4, device=dev)

    def timeit(fn, n=3):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    for impl in (2, 1):
        ms = timeit(lambda: lib.helio_splat_fwd(P(p), B, N, R, 15.0, 15.0, P(img), impl, None))
        print(f"fwd impl {impl}: {ms:.3f} ms  {2*B*N*R*R/ms/1e9:.1f} TFLOP/s (2 FLOP/eval)", flush=True)
    for impl in (2, 1):
        ms = timeit(lambda: lib.helio_splat_bwd(P(p), P(g), B, N, R, 15.0, 15.0, P(mom), impl, None), n=2 if impl == 1 else 3)
        print(f"bwd impl {impl}: {ms:.3f} ms  {4*B*N*R*R/ms/1e9:.1f} TFLOP/s (4 FLOP/eval)", flush=True)
sys.exit(0 if ok else 1)
