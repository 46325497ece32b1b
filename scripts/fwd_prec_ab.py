"""Forward splat: 3xTF32 vs the opt-in f16x3 operand format -- time and error against an fp64 dense evaluation."""
import ctypes as C, sys
sys.path.insert(0, ".")
import torch
from doodle_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
P = lambda t: C.c_void_p(t.data_ptr())
def make_params(B, N, seed=1, sig=(0.8, 1.4)):
    g = torch.Generator(device=dev).manual_seed(seed)
    p = torch.empty(B, N, 4, device=dev)
    p[..., 0] = (torch.rand(B, N, device=dev, generator=g) - 0.5) * 10
    p[..., 1] = (torch.rand(B, N, device=dev, generator=g) - 0.5) * 10
    p[..., 2] = 1.4427 / (2 * (sig[0] + (sig[1] - sig[0]) * torch.rand(B, N, device=dev, generator=g)) ** 2)
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
def dense64(p, R):
    B, N = p.shape[:2]
    xs = torch.linspace(-7.5, 7.5, R, device=dev, dtype=torch.float32).double()
    pd = p.double()
    gx = torch.exp2(-pd[..., 2, None] * (xs[None, None] - pd[..., 0, None]) ** 2) * pd[..., 3, None]
    gy = torch.exp2(-pd[..., 2, None] * (xs[None, None] - pd[..., 1, None]) ** 2)
    return torch.einsum("bni,bnj->bij", gx, gy)
for (B, N, R, sig) in [(4096, 2000, 256, (0.8, 1.4)), (2, 2000, 256, (0.8, 1.4)), (2, 2000, 256, (0.1, 0.3)), (4, 500, 128, (0.3, 2.0)), (3, 300, 64, (0.5, 1.0)),
                       (1024, 5000, 128, (0.8, 1.4)), (16384, 500, 128, (0.8, 1.4)), (1024, 500, 512, (0.8, 1.4)), (1024, 5000, 64, (0.8, 1.4))]:
    p = make_params(B, N, sig=sig)
    out = []
    ref = dense64(p, R) if B <= 4 else None
    for prec in (0, 1):
        assert lib.helio_set_fwd_precision(prec) == 0
        img, ms = run(p, R, 2)
        err = ""
        if ref is not None:
            tol = (img.double() - ref).abs() / (1e-6 + 1e-4 * ref.abs())
            rel = ((img.double() - ref).abs() / ref.abs().clamp_min(1e-30))[ref > 1e-3].max()
            err = f" tol-ratio {float(tol.max()):.4f} max-rel {float(rel):.2e}"
        out.append(f"prec={prec}: {ms*1e3:9.1f} us{err}")
    lib.helio_set_fwd_precision(0)
    print(f"B={B} N={N} R={R} sigma={sig}: " + "   ".join(out), flush=True)
