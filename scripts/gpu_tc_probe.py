"""Quick probe of the tcgen05 forward splat against the SIMT path (run under gpurun with a timeout)."""
import sys, time
sys.path.insert(0, ".")
import torch
from doodle_b200 import functional as Fn, _lib
import ctypes as C
lib = _lib.load()
dev = torch.device("cuda:0")
torch.manual_seed(0)
def run(B, N, R, impl):
    p = torch.empty(B, N, 4, device=dev)
    p[..., 0] = (torch.rand(B, N, device=dev) - 0.5) * 10
    p[..., 1] = (torch.rand(B, N, device=dev) - 0.5) * 10
    p[..., 2] = 1.4427 / (2 * (0.8 + 0.6 * torch.rand(B, N, device=dev)) ** 2)
    p[..., 3] = 0.9 + 0.1 * torch.rand(B, N, device=dev)
    img = torch.full((B, R, R), float("nan"), device=dev)
    rc = lib.helio_splat_fwd(C.c_void_p(p.data_ptr()), B, N, R, 15.0, 15.0, C.c_void_p(img.data_ptr()), impl, None)
    torch.cuda.synchronize()
    assert rc == 0, lib.helio_last_error()
    return p, img
for (B, N, R) in [(1, 32, 256), (2, 40, 256), (3, 300, 256), (2, 64, 128), (5, 100, 200), (300, 64, 256)]:
    torch.manual_seed(1)
    p, a = run(B, N, R, 1)
    torch.manual_seed(1)
    p2, b = run(B, N, R, 2)
    err = ((a - b).abs() / (1e-6 + 1e-4 * a.abs())).max().item()
    print(f"B={B} N={N} R={R}: simt max {a.max().item():.4f} tc max {b.max().item():.4f} nan {torch.isnan(b).sum().item()} tol-ratio {err:.3f}", flush=True)
# timing at the headline shape
B, N, R = 4096, 2000, 256
p = torch.empty(B, N, 4, device=dev)
p[..., 0] = (torch.rand(B, N, device=dev) - 0.5) * 10
p[..., 1] = (torch.rand(B, N, device=dev) - 0.5) * 10
p[..., 2] = 1.4427 / (2 * (0.8 + 0.6 * torch.rand(B, N, device=dev)) ** 2)
p[..., 3] = 1.0
img = torch.empty(B, R, R, device=dev)
for impl in (1, 2):
    for _ in range(2):
        lib.helio_splat_fwd(C.c_void_p(p.data_ptr()), B, N, R, 15.0, 15.0, C.c_void_p(img.data_ptr()), impl, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        lib.helio_splat_fwd(C.c_void_p(p.data_ptr()), B, N, R, 15.0, 15.0, C.c_void_p(img.data_ptr()), impl, None)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"impl {impl}: {ms:.3f} ms  {2*B*N*R*R/ms/1e9:.1f} TFLOP/s (2 FLOP/eval)", flush=True)
