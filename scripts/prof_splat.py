"""Single-kernel driver for ncu: runs helio_splat_fwd / helio_splat_bwd once or twice on a fixed shape."""
import argparse, sys
sys.path.insert(0, ".")
import ctypes as C
import torch
from doodle_b200 import _lib
ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=592); ap.add_argument("--N", type=int, default=2000); ap.add_argument("--R", type=int, default=256)
ap.add_argument("--impl", type=int, default=2); ap.add_argument("--what", default="fwd"); ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--bprec", type=int, default=0, help="1: the f16x3 backward through the debug entry of a -DHELIO_TC_STATS=1 build (HELIO_LIB_PATH)")
a = ap.parse_args()
lib = _lib.load(); dev = torch.device("cuda:0"); torch.manual_seed(0)
B, N, R = a.B, a.N, a.R
p = torch.empty(B, N, 4, device=dev)
p[..., 0] = (torch.rand(B, N, device=dev) - 0.5) * 10
p[..., 1] = (torch.rand(B, N, device=dev) - 0.5) * 10
p[..., 2] = 1.4427 / (2 * (0.8 + 0.6 * torch.rand(B, N, device=dev)) ** 2)
p[..., 3] = 1.0
img = torch.empty(B, R, R, device=dev); g = torch.randn(B, R, R, device=dev); mom = torch.empty(B, N, 4, device=dev)
P = lambda t: C.c_void_p(t.data_ptr())
if a.bprec == 1:
    gmax = g.abs().amax((1, 2)).contiguous()
    lib.helio_debug_splat_bwd_f16.argtypes = [C.c_void_p] * 3 + [C.c_int] * 3 + [C.c_float] * 2 + [C.c_void_p] * 2
    lib.helio_splat_bwd = lambda pp, gg, B_, N_, R_, w_, h_, mm, impl_, st: lib.helio_debug_splat_bwd_f16(pp, gg, P(gmax), B_, N_, R_, w_, h_, mm, st)
for _ in range(a.iters):
    if a.what == "fwd":
        rc = lib.helio_splat_fwd(P(p), B, N, R, 15.0, 15.0, P(img), a.impl, None)
    else:
        rc = lib.helio_splat_bwd(P(p), P(g), B, N, R, 15.0, 15.0, P(mom), a.impl, None)
    assert rc == 0, lib.helio_last_error()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
if a.what == "fwd":
    lib.helio_splat_fwd(P(p), B, N, R, 15.0, 15.0, P(img), a.impl, None)
else:
    lib.helio_splat_bwd(P(p), P(g), B, N, R, 15.0, 15.0, P(mom), a.impl, None)
e1.record(); torch.cuda.synchronize()
out = img if a.what == "fwd" else mom
digest = int(out.view(torch.int32).to(torch.int64).sum().item()) & 0xFFFFFFFFFFFF      # order-independent checksum of the result bits
print(f"{a.what} impl {a.impl} B={B} N={N} R={R}: {e0.elapsed_time(e1):.3f} ms  bits {digest:012x}")
