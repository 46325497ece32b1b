"""Forward splat timing at several shapes (tcgen05 path) + parity vs the CUDA-core path; HELIO_TC_FWD_SPLIT selects the
producer split (read once per process)."""
import ctypes as C, os, sys
sys.path.insert(0, ".")
import torch
from doodle_b200 import _lib
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
def run(p, R, impl, iters=5):
    B, N = p.shape[:2]
    img = torch.empty(B, R, R, device=dev)
    for _ in range(2):
        assert lib.helio_splat_fwd(P(p), B, N, R, 15.0, 15.0, P(img), impl, None) == 0, lib.helio_last_error()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        lib.helio_splat_fwd(P(p), B, N, R, 15.0, 15.0, P(img), impl, None)
    e1.record(); torch.cuda.synchronize()
    return img, e0.elapsed_time(e1) / iters
print("split =", os.environ.get("HELIO_TC_FWD_SPLIT", "auto"))
for (B, N, R) in [(4096, 2000, 256), (1024, 5000, 128), (16384, 500, 128), (1024, 5000, 64), (16384, 50, 64), (1024, 500, 512), (300, 77, 100), (64, 200, 48)]:
    p = make_params(B, N)
    img, ms = run(p, R, 2)
    err = ""
    if B * N * R * R < 3e11:
        ref, _ = run(p, R, 1, iters=1)
        err = f" tol-ratio {float(((img - ref).abs() / (1e-6 + 1e-4 * ref.abs())).max()):.3f}"
    print(f"B={B} N={N} R={R}: {ms*1e3:9.1f} us{err}", flush=True)
