#!/usr/bin/env python3
"""Per-role cycle accounting of the tcgen05 splat kernels (debug build, -DHELIO_TC_STATS=1):

    make -C doodle_b200/csrc dbg DBGFLAGS="-DHELIO_TC_STATS=1"
    HELIO_LIB_PATH=doodle_b200/libhelio_dbg.so python scripts/tc_stats.py [--what fwd|bwd] [--B 4096] [--N 2000] [--R 256] [--prec 0|1]

Every warp of every CTA records the cycles it spent inside its role loop and the part of them it was blocked in each of
its waits.  Printed per role, averaged over CTAs: which side of the pipeline waits for which.
  producers : commit = fence.proxy.async + arrive (drain of the operand stores), acquire = wait for the slot (MMAs retired)
  MMA warp  : tempty = wait for the epilogue to free the accumulator, full = wait for the producers
  epilogue  : tfull  = wait for the tile's MMAs
"""
import argparse, ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from doodle_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=4096); ap.add_argument("--N", type=int, default=2000); ap.add_argument("--R", type=int, default=256)
ap.add_argument("--what", default="both"); ap.add_argument("--prec", type=int, default=0); ap.add_argument("--pair", type=int, default=0)
ap.add_argument("--bprec", type=int, default=0, help="backward operand format: 0 = 3xTF32, 1 = f16x3 K=64 (debug entry)")
a = ap.parse_args()
lib = _lib.load()
assert hasattr(lib, "helio_debug_tc_stats"), "load a -DHELIO_TC_STATS=1 build through HELIO_LIB_PATH"
lib.helio_debug_tc_stats.argtypes = [C.c_void_p, C.c_int]
dev = torch.device("cuda:0"); torch.manual_seed(0)
B, N, R = a.B, a.N, a.R
p = torch.empty(B, N, 4, device=dev)
p[..., 0] = (torch.rand(B, N, device=dev) - 0.5) * 10
p[..., 1] = (torch.rand(B, N, device=dev) - 0.5) * 10
p[..., 2] = 1.4427 / (2 * (0.8 + 0.6 * torch.rand(B, N, device=dev)) ** 2)
p[..., 3] = 1.0
img = torch.empty(B, R, R, device=dev); g = torch.randn(B, R, R, device=dev); mom = torch.empty(B, N, 4, device=dev)
P = lambda t: C.c_void_p(t.data_ptr())
lib.helio_set_fwd_precision(a.prec); lib.helio_set_tc_pair_mode(a.pair)


gmax = g.abs().amax((1, 2)).contiguous()


def run(what):
    if what == "fwd":
        f = lambda: lib.helio_splat_fwd(P(p), B, N, R, 15.0, 15.0, P(img), 2, None)
    elif a.bprec == 1:
        lib.helio_debug_splat_bwd_f16.argtypes = [C.c_void_p] * 3 + [C.c_int] * 3 + [C.c_float] * 2 + [C.c_void_p] * 2
        f = lambda: lib.helio_debug_splat_bwd_f16(P(p), P(g), P(gmax), B, N, R, 15.0, 15.0, P(mom), None)
    else:
        f = lambda: lib.helio_splat_bwd(P(p), P(g), B, N, R, 15.0, 15.0, P(mom), 2, None)
    for _ in range(2):
        assert f() == 0, lib.helio_last_error()
    torch.cuda.synchronize()
    lib.helio_debug_tc_stats(None, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record(); torch.cuda.synchronize()
    st = np.zeros((160, 24, 4), np.uint64)
    lib.helio_debug_tc_stats(st.ctypes.data, 0)
    st = st.astype(np.float64)
    used = st[:, :, 0].sum(1) > 0
    ncta = int(used.sum())
    print(f"== {what}: B={B} N={N} R={R} prec={a.prec} bprec={a.bprec}: {e0.elapsed_time(e1):.3f} ms, {ncta} CTAs reporting")
    tot = st[used]
    warps = [w for w in range(24) if tot[:, w, 0].max() > 0]
    kernel_cycles = tot[:, :, 0].max()
    print(f"   longest role loop: {kernel_cycles/1e6:.3f} Mcycles")
    for w in warps:
        t = tot[:, w, :]
        live = t[:, 0] > 0
        t = t[live]
        print(f"   warp {w:2d}: loop {t[:,0].mean()/1e6:7.3f} Mcyc | wait1 {100*t[:,1].sum()/t[:,0].sum():5.1f}% | wait2 {100*t[:,2].sum()/t[:,0].sum():5.1f}% "
              f"| busy {100*(1-(t[:,1].sum()+t[:,2].sum())/t[:,0].sum()):5.1f}%  (CTAs {int(live.sum())})")


for what in (("fwd", "bwd") if a.what == "both" else (a.what,)):
    run(what)
print("wait1/wait2: producers = commit/acquire; MMA warp = tempty/full; epilogue = tfull/-")
