#!/usr/bin/env python3
"""bench.py -- headline benchmark of the DOODLE flux-renderer hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one ``HelioEnv.step(action)`` (noisy render forward, the four losses, the per-step NaN/Inf
check) followed by ``(mse + dist + bound + alignment_loss).backward()`` down to ``action.grad`` on the
BASELINE.json headline configuration: N=2000 heliostats, 256x256 receiver, B=4096 suns per GPU
(weak scaling: the sun batch is sharded, heliostat geometry replicated, one 4-float all-reduce per
step).  The environment runs with its product defaults: the target image of the error-free field is
cached (exact: it depends on the sun positions only; the reference re-renders the identical image every
step, test_environment.py:429-435) -- the line's ``uncached`` block is the same step with the target
re-rendered every step (``--no-cache-target`` makes that the headline).  ``value`` = heliostat*pixel
evals/s = (B_global * N * R^2) / step time, inputs resident in HBM; ``e2e`` = the same with the action in
pinned host memory (H2D inside the timed region) and the action gradient + metrics read back (D2H).
With N > 1 ranks the line also carries ``strong_scaling``: the same step at a GLOBAL batch of 4096 suns
(4096/N per GPU).

``--impl reference`` times the oracle port of the reference's CPU algorithm (oracle/helio_oracle.py,
numpy, all host threads) on a bounded sample of the same workload; the reference itself is pure
Python that cannot travel to the GPU box.  One JSON line on stdout either way.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(name="large_field N=2000 R=256 B=4096/GPU (BASELINE.json configs[3])", N=2000, R=256, B=4096,
                sigma_scale=0.01, error_scale_mrad=90.0)
METRIC = "heliostat_pixel_evals_per_s (HelioEnv.step fwd + losses + backward to action.grad)"
UNIT = "evals/s"
FLOP_PER_EVAL = {"splat_fwd": 2.0, "splat_bwd": 4.0}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--N", type=int, default=WORKLOAD["N"])
    ap.add_argument("--R", type=int, default=WORKLOAD["R"])
    ap.add_argument("--B", type=int, default=WORKLOAD["B"], help="suns per GPU")
    ap.add_argument("--splat", default="auto", choices=["auto", "simt", "tc"])
    ap.add_argument("--no-cache-target", action="store_true",
                    help="headline = target re-rendered every step as the reference does (default: exact target cache, the product default)")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the torch-eager dense comparator on the same GPU")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling block (N > 1)")
    ap.add_argument("--no-both-3xtf32", action="store_true", help="skip the side measurement with both contractions in 3xTF32")
    ap.add_argument("--e2e-serial", action="store_true", help="e2e arm with caller-side copies instead of the host-action step")
    ap.add_argument("--host-chunks", type=int, default=0, help="slices of the sun batch the host-action step overlaps its copies in (0 = the env's default)")
    ap.add_argument("--e2e-only", action="store_true", help="tuning aid: print the e2e arm's ms per step and exit")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-culled", action="store_true", help="skip the opt-in footprint-culling side measurement")
    ap.add_argument("--no-small-field", action="store_true", help="skip the N=50, R=128, B=25 env-steps/s side measurement")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="budget for the cpu_baseline sample")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# synthetic workload: identical on CPU and GPU
# ---------------------------------------------------------------------------------------------
def make_inputs(N, B, seed=42, rank=0):
    import torch
    g = torch.Generator().manual_seed(seed)
    helio = torch.rand(N, 3, generator=g) * 10 + 80          # train_with_env.py:227
    helio[:, 2] = 0
    targ_pos = torch.tensor([0., -5., 0.])
    targ_norm = torch.tensor([0., 1., 0.])
    area = (15., 15.)
    g2 = torch.Generator().manual_seed(seed + 1 + rank)
    return helio, targ_pos, targ_norm, area, g2


def sample_suns(B, gen):
    """B directions in a 2-degree cone about az=el=45 deg, |z|, radius hypot(1e4,1e4) (test_environment.py:42-88,293,324):
    the product's own sampler, seeded from ``gen`` without disturbing the global generator."""
    import torch
    from doodle_b200.env import azimuth_elevation_to_primary_direction, sample_cone_directions
    seed = int(torch.randint(0, 2 ** 31 - 1, (1,), generator=gen).item())
    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(seed)
        d = sample_cone_directions(B, azimuth_elevation_to_primary_direction(45.0, 45.0), 2.0, force_upper_hemisphere=True)
    return d * math.hypot(10000, 10000)


# ---------------------------------------------------------------------------------------------
# CPU arm: oracle port of the reference algorithm
# ---------------------------------------------------------------------------------------------
def cpu_step_factory(N, R, n_suns, threads):
    import numpy as np
    import torch
    from oracle import helio_oracle as orc
    helio, targ_pos, targ_norm, area, gen = make_inputs(N, n_suns)
    sun = sample_suns(n_suns, gen).numpy()
    helio_n = helio.numpy()
    ideal = orc.calculate_ideal_normals(sun, helio_n, targ_pos.numpy())
    act = ideal + 0.01 * torch.randn(ideal.shape, generator=gen).numpy()
    act = (act / np.linalg.norm(act, axis=2, keepdims=True)).astype(np.float32)
    errs = (torch.randn(n_suns, N, 2, generator=gen) * WORKLOAD["error_scale_mrad"]).numpy()
    dmaps = torch.rand(n_suns, R, R, generator=gen).numpy() * 50

    def step():
        return orc.env_step(sun, act, errs, helio_n, targ_pos.numpy(), targ_norm.numpy(), area, R,
                            WORKLOAD["sigma_scale"], dmaps, threads=threads)
    return step


def run_cpu_sample(N, R, budget_s, steps=1, warmup=0):
    """Time the oracle port on a bounded sample of the workload; returns (dict for `cpu_baseline`, seconds per step).

    The port parallelises over suns, so the sample is one sun per host thread (more when the budget allows) over the
    first N_s heliostats of the field, N_s sized so that (steps + warmup) steps fit `budget_s`.  The dense algorithm is
    linear in the heliostat count, so evals/s of the sample is the rate of the full workload."""
    threads = os.cpu_count() or 1
    n_cal = min(N, 64)
    cal = cpu_step_factory(n_cal, R, 1, 1)
    cal()                                            # first call pays imports / page faults
    t0 = time.perf_counter()
    cal()
    t_unit = (time.perf_counter() - t0) / n_cal      # seconds per (sun, heliostat) on one thread
    per_step_budget = budget_s / max(steps + warmup, 1)
    n_s = int(min(N, max(32, per_step_budget / max(t_unit, 1e-9))))
    n_suns = threads
    if n_s == N:
        n_suns = threads * int(max(1, min(4, per_step_budget / max(t_unit * N, 1e-9))))
    step = cpu_step_factory(n_s, R, n_suns, threads)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(max(steps, 1)):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    evals = float(n_suns) * n_s * R * R
    return dict(value=evals / dt, unit=UNIT, cores=threads, kind="port",
                sample=f"oracle/helio_oracle.env_step (numpy port of the reference's dense algorithm, fp32): "
                       f"{n_suns} suns x {n_s} of N={N} heliostats x R={R}, step fwd+target+losses+bwd, {dt:.2f} s/step on {threads} threads"), dt


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps, warmup = max(args.steps, 1), max(args.warmup, 0)
    cb, dt = run_cpu_sample(args.N, args.R, budget_s=150.0, steps=steps, warmup=warmup)
    line = dict(impl="reference", metric=METRIC, value=cb["value"], unit=UNIT, n_gpus=args.gpus, steps=steps, warmup=warmup,
                ms_per_step=dt * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=WORKLOAD["name"], N=args.N, R=args.R, B_per_gpu=args.B, note="bounded sample, see cpu_baseline.sample"),
                cpu_baseline=cb,
                e2e=dict(value=cb["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle reasons DURING the timed region: an NVML polling thread (10 ms period), or
    `nvidia-smi -lms` when pynvml is unavailable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        import threading
        self.samples, self.reasons, self.power, self.mx = [], set(), [], None
        self.p = self.f = self.thread = None
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].strip().isdigit() else gpu_index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            names = {"hw_slowdown": pynvml.nvmlClocksThrottleReasonHwSlowdown,
                     "hw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonSwThermalSlowdown,
                     "sw_power_cap": pynvml.nvmlClocksThrottleReasonSwPowerCap}

            def poll():
                while not self._stop.is_set():
                    try:
                        self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        self.power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
                        r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        for k, bit in names.items():
                            if r & bit:
                                self.reasons.add(k)
                    except Exception:
                        pass
                    self._stop.wait(0.01)
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            try:
                self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                           "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
            except Exception:
                self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.thread is not None:
            self._stop.set()
            self.thread.join(timeout=2)
            if self.samples:
                sm = sorted(self.samples)
                out.update(sm_mhz=sm[len(sm) // 2], sm_min_mhz=sm[0], sm_max_mhz=self.mx, reasons=sorted(self.reasons),
                           samples=len(sm), power_w_max=max(self.power) if self.power else None, source="nvml")
            return out
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.f.read().strip().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_min_mhz=sm[0], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power), source="nvidia-smi")
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def measure_tf32_peak(torch):
    """cuBLAS TF32 GEMM burst rate on this box (denominator of the 3xTF32 roofline = this / 3)."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device="cuda"); b = torch.randn(n, n, device="cuda")
        for _ in range(3):
            a @ b
        best = float("inf")
        for _ in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def small_field_numbers(torch, dev, iters=300):
    """BASELINE.json configs[1]: HelioEnv reset/step, N=50, 128x128, B=25, new errors every reset, all four losses +
    the caller's own loss.backward().  Launch/host-latency bound, so reported as env-steps/s through the PUBLIC API with
    its defaults (graph="auto": transparent CUDA-graph replay inside env.step, deferred finite check), the same with
    graph=False (eager fused step, one sync per step for the NaN/Inf asserts) and the explicit one-graph GraphedStep."""
    from doodle_b200 import GraphedStep, HelioEnv
    N, R, B = 50, 128, 25
    helio, targ_pos, targ_norm, area, _ = make_inputs(N, B)

    def make_env(**kw):
        torch.manual_seed(7)
        env = HelioEnv(heliostat_pos=helio.to(dev), targ_pos=targ_pos.to(dev), targ_area=area, targ_norm=targ_norm.to(dev), sigma_scale=0.1,
                       error_scale_mrad=90.0, resolution=R, batch_size=B, device=str(dev), new_errors_every_reset=True, **kw)
        env.reset()
        return env

    def clock(fn, n):
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    out = {}
    grads = {}
    for name, kw in (("default", {}), ("eager", dict(graph=False))):
        env = make_env(**kw)
        a0 = env.ideal_normals.flatten(1).clone()

        def step(env=env, a0=a0):
            a = a0.detach().requires_grad_(True)
            obs, m, mon = env.step(a)
            (m["mse"] + m["dist"] + m["bound"] + m["alignment_loss"]).backward()
            return a.grad
        out[name] = clock(step, iters)
        grads[name] = step().clone()
        if name == "default":
            replaying = env._step_graph is not None
            gs = GraphedStep(env)
            out["explicit"] = clock(lambda: gs(a0), iters)
            kernels = gs.helio_kernels_per_replay
    ms = out["default"]
    return dict(workload="HelioEnv.step + caller's loss.backward(), N=50 R=128 B=25 (BASELINE.json configs[1])",
                env_steps_per_s=1e3 / ms, us_per_step=ms * 1e3, public_api_replays_graphs=bool(replaying),
                us_per_step_eager=out["eager"] * 1e3, env_steps_per_s_eager=1e3 / out["eager"],
                us_per_step_graphed=out["explicit"] * 1e3, env_steps_per_s_graphed=1e3 / out["explicit"],
                bit_equal_to_eager=bool(torch.equal(grads["default"], grads["eager"])),
                evals_per_s=B * N * R * R / (ms * 1e-3), helio_kernels_per_step=kernels,
                note="default = env.step through the public API (graph='auto': graph replay inside step, autograd-connected metrics, "
                     "finite check deferred to the next step); eager = graph=False (1 sync per step); graphed = explicit GraphedStep "
                     "(step + objective + backward in one graph)")


def gpu_eager_numbers(torch, dev, N, R, budget_s=20.0):
    """The reference's own GPU path is stock ATen eager on the dense algorithm: time its restatement
    (oracle/helio_torch_eager.py, pinned on the reference's fixtures) on this GPU, fwd + target + losses + backward,
    on as many suns as the dense intermediates allow (~50-60 B per heliostat-pixel held for autograd)."""
    from oracle import helio_torch_eager as te
    helio, targ_pos, targ_norm, area, gen = make_inputs(N, 1)
    free = torch.cuda.mem_get_info(dev)[0]
    per_sun = float(N) * R * R * 90.0                       # bytes: autograd-held + transient [M,R,R,3] temporaries, with headroom
    chunk = int(max(1, min(64, free * 0.6 / per_sun)))
    sun = sample_suns(chunk, gen).to(dev)
    helio_d, tp, tn = helio.to(dev), targ_pos.to(dev), targ_norm.to(dev)
    ideal = te.ideal_normals(sun, helio_d, tp)
    act = torch.nn.functional.normalize(ideal + 0.01 * torch.randn(ideal.shape, generator=gen).to(dev), dim=2)
    errs = (torch.randn(chunk, N, 2, generator=gen) * WORKLOAD["error_scale_mrad"]).to(dev)
    dmaps = (torch.rand(chunk, R, R, generator=gen) * 50).to(dev)

    def step():
        return te.chunked_step_and_backward(sun, act, errs, helio_d, tp, tn, area, R, WORKLOAD["sigma_scale"], dmaps, chunk=chunk)
    try:
        step()
    except torch.cuda.OutOfMemoryError:
        torch.cuda.empty_cache()
        chunk = max(1, chunk // 2)
        sun, act, errs, dmaps = sun[:chunk], act[:chunk], errs[:chunk], dmaps[:chunk]
        step()
    torch.cuda.synchronize()
    times = []
    t_end = time.perf_counter() + budget_s
    while len(times) < 5 and (time.perf_counter() < t_end or not times):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    peak = torch.cuda.max_memory_allocated(dev)
    del sun, act, errs, dmaps
    torch.cuda.empty_cache()
    ms = min(times)
    return dict(value=float(chunk) * N * R * R / (ms * 1e-3), unit=UNIT, chunk_suns=chunk, ms_per_chunk=ms, repeats=len(times),
                peak_mem_gb=peak / 2 ** 30, kind="torch-eager dense restatement of the reference's op chain (oracle/helio_torch_eager.py), "
                "stock ATen kernels + autograd on this GPU, fp32; step = noisy render + target render + 4 losses + backward",
                note="the reference holds ~50 B per heliostat-pixel for autograd, so the batch is processed chunk_suns at a time; "
                     "evals/s of a chunk is the rate of the full workload (independent suns)")


def _snapshot_step(torch, env, action0):
    """(image, action gradient of mse + bound + alignment, action gradient of the full objective) of one env step on the bench
    inputs.  The distance-weighted L1 term is left out of the first gradient: its image gradient is sign(diff) * dmap, a step
    function of the image, so two forward passes that differ in the last bits flip it on pixels where |diff| is at rounding level
    (the reference's own fp32 does, under any reordering) -- that measures the kink of |x|, not the contraction's accuracy."""
    a = action0.detach().clone().requires_grad_(True)
    obs, m, _ = env.step(a)
    (m["mse"] + m["bound"] + m["alignment_loss"]).backward(retain_graph=True)
    g_smooth = a.grad.detach().reshape(-1, 3).clone()
    a.grad = None
    (m["mse"] + m["dist"] + m["bound"] + m["alignment_loss"]).backward()
    return obs["img"].detach().clone(), g_smooth, a.grad.detach().reshape(-1, 3).clone()


def _compare_snapshots(torch, got, ref):
    """How far the step in the default operand formats (f16x3) is from the same step in 3xTF32, in units of BASELINE.json's
    tolerances: images 1e-4 relative + 1e-6 absolute; action gradients 1e-3 relative, elementwise as tests/conftest.py checks
    them (1e-3 |ref| + 1e-3 median|ref| + 1e-5 of the heliostat's gradient norm).  < 1 = inside the tolerance."""
    (img, g, gf), (img_r, g_r, gf_r) = got, ref
    img_ratio = float(((img - img_r).abs() / (1e-6 + 1e-4 * img_r.abs())).max())
    live = g_r.abs()[g_r.abs() > 1e-30]
    med = float(live.median()) if live.numel() else 0.0
    tol = 1e-3 * g_r.abs() + 1e-3 * med + 1e-5 * g_r.abs().amax(dim=1, keepdim=True)
    g_ratio = float(((g - g_r).abs() / tol.clamp_min(1e-38)).max())
    return dict(image_max_diff_over_tolerance=img_ratio, gradient_max_diff_over_tolerance=g_ratio,
                image_max_rel_diff=float(((img - img_r).abs().max()) / img_r.abs().max().clamp_min(1e-30)),
                gradient_max_norm_rel_diff=float((g - g_r).abs().max() / g_r.abs().max().clamp_min(1e-30)),
                full_objective_gradient_max_norm_rel_diff=float((gf - gf_r).abs().max() / gf_r.abs().max().clamp_min(1e-30)),
                note="same inputs, same cached target; tolerances of BASELINE.json (images rtol 1e-4 / atol 1e-6, gradients 1e-3 "
                     "relative, elementwise); values < 1 are inside the tolerance.  The elementwise gradient figure is for mse + bound "
                     "+ alignment; the distance-weighted L1 term's image gradient is sign(diff) * dmap, which flips on pixels whose "
                     "|diff| is at rounding level whenever the forward changes in its last bits (in any fp32 implementation), so the "
                     "full objective is reported in the max norm only")


def _leave_process_group(torch, dist, envs):
    """Release captured graphs (a sharded small field's graphs hold NCCL kernels), then destroy the process group under a
    watchdog: a hung teardown must not cost GPU time."""
    import threading
    for e in envs:
        try:
            if e is not None:
                e.close()
        except Exception:
            pass
    torch.cuda.synchronize()
    sys.stdout.flush()
    t = threading.Timer(30.0, lambda: os._exit(0))
    t.daemon = True
    t.start()
    dist.destroy_process_group()
    t.cancel()


def main_ours(args):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a B200; doodle_b200 has no CPU fallback")
    from doodle_b200 import HelioEnv, functional as Fn, _lib
    from doodle_b200.dist import make_sharded_env

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    N, R, B = args.N, args.R, args.B
    cached = not args.no_cache_target
    impl = dict(auto=_lib.SPLAT_AUTO, simt=_lib.SPLAT_SIMT, tc=_lib.SPLAT_TC)[args.splat]

    def build_env(B_local, cache_target=cached):
        helio, targ_pos, targ_norm, area, gen = make_inputs(N, B_local, rank=rank)
        kw = dict(heliostat_pos=helio.to(dev), targ_pos=targ_pos.to(dev), targ_area=area, targ_norm=targ_norm.to(dev),
                  sigma_scale=WORKLOAD["sigma_scale"], error_scale_mrad=WORKLOAD["error_scale_mrad"], initial_action_noise=0.0,
                  resolution=R, device=str(dev), new_errors_every_reset=True, cache_target=cache_target)   # check_finite: product default (on)
        if world > 1:
            env = make_sharded_env(HelioEnv, global_batch_size=B_local * world, seed=42, **kw)   # same seed on every rank: global draws, sliced
        else:
            torch.manual_seed(42)
            env = HelioEnv(batch_size=B_local, **kw)
        env.noisy_field.splat_impl = impl
        env.ref_field.splat_impl = impl
        env.reset()
        torch.manual_seed(1000 + rank)
        a0 = env.noisy_field.initial_action.detach().clone().view(B_local, N, 3)
        a0 = a0 + 0.01 * torch.randn_like(a0)
        return env, (a0 / a0.norm(dim=2, keepdim=True)).contiguous()

    env, action0 = build_env(B)

    def one_step(action, env=None):
        env = env or ENV[0]
        action = action.detach().requires_grad_(True)
        obs, metrics, monitor = env.step(action)
        loss = metrics["mse"] + metrics["dist"] + metrics["bound"] + metrics["alignment_loss"]
        loss.backward()
        return action.grad, metrics
    ENV = [env]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    steps, warmup = max(args.steps, 1), max(args.warmup, 3)
    # ---- device-resident arm --------------------------------------------------------------
    for _ in range(warmup):
        one_step(action0)
    Fn.reset_profile(True)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = Fn.launch_count()
    ms_total = timed(lambda: one_step(action0), steps)
    launches = Fn.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    try:
        mhz_fwd, mhz_bwd = Fn.tc_clock_mhz()          # SM clock the tcgen05 kernels HELD (measured inside the kernels)
    except Exception:
        mhz_fwd = mhz_bwd = 0.0
    kprof = Fn.collect_profile()
    Fn.reset_profile(False)
    ms_step = ms_total / steps
    evals_step = float(B) * world * N * R * R

    # ---- end-to-end arm: host action in, host gradient + metrics out ------------------------
    h_action = action0.cpu().pin_memory()
    h_grad = torch.empty_like(h_action).pin_memory()
    h_metrics = torch.empty(4, dtype=torch.float32).pin_memory()

    def e2e_step():
        if args.e2e_serial:                            # plain caller-side copies around a device step
            a = h_action.to(dev, non_blocking=True)
            g, m = one_step(a)
            h_grad.copy_(g, non_blocking=True)
        else:                                          # public API with a host action: env.step overlaps the copies
            g, m = one_step(h_action)                  # g = action.grad, already in (pinned) host memory
            assert g.device.type == "cpu"
        h_metrics.copy_(torch.stack([m["mse"], m["dist"], m["bound"], m["alignment_loss"]]).detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the caller reads the result every step

    if args.host_chunks > 0:
        env.host_chunks = args.host_chunks
    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, steps) / steps
    if args.e2e_only:
        if rank == 0:
            print(json.dumps(dict(e2e_ms_per_step=ms_e2e, host_chunks=env._host_chunks(B), device_ms_per_step=ms_step)))
        if world > 1:
            _leave_process_group(torch, dist, [ENV[0]])
        return 0

    # ---- host link while every rank copies at once (explains the e2e scaling) ------------------------------------
    def link_gbs(dst, src, reps=5):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([src.numel() * 4 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return float(t.item())
    d_tmp = torch.empty_like(action0)
    h2d_gbs, d2h_gbs = link_gbs(d_tmp, h_action), link_gbs(h_grad, d_tmp)
    del d_tmp

    # ---- the other target mode on the same box (cached <-> re-rendered every step) -----------------------------------
    other = None
    try:
        env.cache_target = not cached
        env._target_cache = None
        for _ in range(3):
            one_step(action0)
        Fn.reset_profile(True)
        ms_o = timed(lambda: one_step(action0), steps) / steps
        ko = Fn.collect_profile()
        Fn.reset_profile(False)
        other = dict(target_cached=not cached, ms_per_step=ms_o, value=evals_step / (ms_o * 1e-3), unit=UNIT,
                     kernels_ms={k: round(v["avg_ms"], 4) for k, v in sorted(ko.items())},
                     launches_per_step={k: v["n"] / steps for k, v in sorted(ko.items())})
    except Exception as e:
        other = dict(error=repr(e))
    finally:
        env.cache_target = cached
        env._target_cache = None

    # ---- strong scaling: the same step at a GLOBAL batch of B suns (B / world per GPU) -------------------------------
    strong = None
    if world > 1 and not args.no_strong and B % world == 0:
        try:
            del env
            ENV[0] = None
            torch.cuda.empty_cache()
            env_s, a_s = build_env(B // world)
            ENV[0] = env_s
            for _ in range(warmup):
                one_step(a_s)
            ms_s = timed(lambda: one_step(a_s), steps) / steps
            strong = dict(global_batch=B, suns_per_gpu=B // world, ms_per_step=ms_s, value=float(B) * N * R * R / (ms_s * 1e-3), unit=UNIT,
                          scaling="strong", target_cached=cached,
                          note="same step, total work fixed at the 1-GPU batch; efficiency = value / (n_gpus x the 1-GPU line's value)")
            env = env_s
        except Exception as e:
            strong = dict(error=repr(e))

    if rank != 0:
        if world > 1:
            _leave_process_group(torch, dist, [ENV[0]])
        return 0

    # ---- roofline of the dominant kernel ------------------------------------------------------
    # peak = fp32-accurate tensor rate = TF32 dense / 3 (3xTF32), TF32 dense = 1/2 of the bf16 dense rate in
    # MEASURED_PEAKS.json.  Reported against the sustained figure (the kernel is timed inside a long step) AND the burst
    # figure, plus the physical tensor-pipe utilisation at the clock the kernel held.
    peaks, peak_src = {}, "of fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
        peak_src = "of measured"
    except Exception:
        pass
    bf16 = float(peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops") or 1400.0)
    bf16_burst = float(peaks.get("bf16_tflops") or bf16)
    tf32 = measure_tf32_peak(torch)
    dom = max(("splat_fwd", "splat_bwd"), key=lambda k: kprof.get(k, {}).get("total_ms", 0.0))
    # operand format of the dominant kernel: "3xtf32" (three kind::tf32 MMAs per MAC, peak = TF32 dense / 3 = bf16 / 6) or
    # "f16x3" (three kind::f16 MMAs per MAC at twice the rate, peak = bf16 dense / 3); same 22-bit operand accuracy
    fwd_mode = int(os.environ.get("HELIO_FWD_PREC", "2"))
    bwd_mode = int(os.environ.get("HELIO_BWD_PREC", str(_lib.BWD_PREC_DEFAULT)))
    fmt = {"splat_fwd": "f16x3" if fwd_mode in (1, 2) else "3xtf32",
           "splat_bwd": "f16x3" if bwd_mode == 1 else "3xtf32"}
    per_mac = 3.0 if fmt[dom] == "f16x3" else 6.0           # bf16-dense FLOP spent per algorithmic FLOP
    peak = bf16 / per_mac
    d = kprof.get(dom, {})
    avg_ms = d.get("avg_ms") or float("nan")
    flops_launch = FLOP_PER_EVAL[dom] * float(B) * N * R * R
    achieved = flops_launch / (avg_ms * 1e-3) / 1e12
    traffic = None
    try:   # dram bytes per launch of this kernel at this shape, from the committed ncu --set full capture
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f).get(f"{dom}:N{N}:R{R}:B{B}")
        traffic = t["dram_bytes"] if t else None
    except Exception:
        pass
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    n_sms = torch.cuda.get_device_properties(dev).multi_processor_count
    # executed tensor work: 3 tcgen05.mma per algorithmic MAC (hi*hi, hi*lo, lo*hi); the forward pads K = heliostats to 32 per
    # stage, the backward pads the heliostat rows of a tile to 128 per CTA (256 per CTA pair at R > 128)
    pad = ((N + 31) // 32 * 32) / float(N) if dom == "splat_fwd" else ((N + 255) // 256 * 256 if R > 128 else (N + 127) // 128 * 128) / float(N)
    executed = 3.0 * flops_launch * pad
    pipe_rate = 8192.0 if fmt[dom] == "f16x3" else 4096.0   # tensor FLOP / clk / SM of the MMA kind in use
    pipe_frac = executed / (avg_ms * 1e-3) / (n_sms * pipe_rate * sm_mhz * 1e6)
    mhz_dom = mhz_fwd if dom == "splat_fwd" else mhz_bwd
    pipe_frac_held = executed / (avg_ms * 1e-3) / (n_sms * pipe_rate * mhz_dom * 1e6) if mhz_dom else None
    roofline = dict(bound="tensor", kernel=dom, achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak, traffic=traffic,
                    operand_format=fmt,
                    peak_source=f"{peak_src}: bf16 {bf16:.0f} TFLOP/s sustained / " + ("3 (f16x3: three kind::f16 MMAs per MAC)" if fmt[dom] == "f16x3"
                                                                                       else "2 (TF32) / 3 (3xTF32)"),
                    frac_of_burst_peak=achieved / (bf16_burst / per_mac), burst_peak=bf16_burst / per_mac,
                    tensor_pipe_frac=pipe_frac,
                    sm_mhz_held_in_kernel=dict(splat_fwd=mhz_fwd, splat_bwd=mhz_bwd,
                                               note="cycles / nanoseconds measured INSIDE the tcgen05 kernels (helio_tc_clock_mhz): B200 power-throttles under "
                                                    "sustained tensor load while NVML keeps reporting the application clock"),
                    tensor_pipe_frac_at_held_clock=pipe_frac_held,
                    tensor_pipe_note=f"executed tcgen05 FLOP (3 x algorithmic x padding) / ({n_sms} SMs x {pipe_rate:.0f} FLOP/clk ({fmt[dom]}) x {sm_mhz:.0f} MHz median under load)",
                    cublas_tf32_inrun_tflops=tf32, frac_of_cublas_tf32_over_3=achieved / (tf32 / 3.0),
                    frac_of_nominal_peak=achieved / (2250.0 / per_mac),   # 2.25 PFLOP/s bf16 nominal dense
                    avg_launch_ms=avg_ms, share_of_step=d.get("total_ms", 0.0) / ms_total if ms_total else None,
                    kernels_ms={k: round(v["avg_ms"], 4) for k, v in sorted(kprof.items())},
                    launches_per_step={k: v["n"] / steps for k, v in sorted(kprof.items())},
                    note=f"algorithmic {FLOP_PER_EVAL[dom]:.0f} FLOP/eval x {B*N*R*R:.3e} evals per launch (SURVEY 8d); the other tensor kernel: "
                         + ", ".join(f"{k} {FLOP_PER_EVAL[k] * float(B) * N * R * R / (kprof[k]['avg_ms'] * 1e-3) / 1e12:.0f} TFLOP/s"
                                     for k in ("splat_fwd", "splat_bwd") if k != dom and k in kprof)
                         + "; traffic = dram read+write bytes per launch from profiles/ (ncu --set full); frac > 1 is possible against the sustained "
                           "figure because MEASURED_PEAKS.json took it power-throttled (1335 MHz): frac_of_burst_peak and "
                           "tensor_pipe_frac_at_held_clock are the informative ones")

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:      # reported at N=1 only
        cpu_baseline, _ = run_cpu_sample(N, R, budget_s=args.cpu_seconds)
    gpu_eager = None
    if not args.no_gpu_eager and world == 1:
        try:
            del env
            ENV[0] = None
            torch.cuda.empty_cache()
            gpu_eager = gpu_eager_numbers(torch, dev, N, R)
        except Exception as e:
            gpu_eager = dict(error=repr(e))
        env, action0 = build_env(B)
        ENV[0] = env
    # ---- opt-in footprint culling on the same workload (secondary; the headline above is the dense evaluation) ----
    culled = None
    if not args.no_culled and world == 1:          # single-GPU side measurement (the other ranks have left by now)
        try:
            env.cull = True
            for _ in range(3):
                one_step(action0)
            Fn.reset_profile(True)
            ms_c = timed(lambda: one_step(action0), steps) / steps
            kc = Fn.collect_profile()
            Fn.reset_profile(False)
            _g, _m = one_step(action0)                 # keep this step's graph alive: it owns the culled lists
            kept = Fn.last_cull_kept_fraction()
            del _g, _m
            culled = dict(ms_per_step=ms_c, dense_equivalent_evals_per_s=evals_step / (ms_c * 1e-3), kept_fraction=kept,
                          kernels_ms={k: round(v["avg_ms"], 4) for k, v in sorted(kc.items())},
                          note="HelioEnv(cull=True): heliostats whose footprint is below 2^-40 of its peak on every pixel are compacted "
                               "away per sun before K2/K3 (helio_cull); same images / gradients within 1e-5")
        except Exception as e:
            culled = dict(error=repr(e))
        finally:
            env.cull = False
    # ---- both contractions in 3xTF32 (the format BASELINE.json names; the backward's default is f16x3) ----
    tf32x3 = None
    if not args.no_both_3xtf32 and world == 1:
        try:
            _lib.check(_lib.load().helio_set_bwd_precision(0), "helio_set_bwd_precision")
            _lib.check(_lib.load().helio_set_fwd_precision(0), "helio_set_fwd_precision")
            for _ in range(3):
                one_step(action0)
            Fn.reset_profile(True)
            ms_t = timed(lambda: one_step(action0), steps) / steps
            kt = Fn.collect_profile()
            Fn.reset_profile(False)
            tf32x3 = dict(ms_per_step=ms_t, evals_per_s=evals_step / (ms_t * 1e-3), kernels_ms={k: round(v["avg_ms"], 4) for k, v in sorted(kt.items())},
                          note="helio_set_fwd_precision(0) + helio_set_bwd_precision(0): both contractions in 3xTF32, the format BASELINE.json "
                               "names (the defaults are the f16x3 formats: same 22-bit operand accuracy, half the tensor work); same "
                               "results within the parity tolerances")
            snap_t = _snapshot_step(torch, ENV[0], action0)            # image + action gradient of one step in 3xTF32
        except Exception as e:
            tf32x3 = dict(error=repr(e))
            snap_t = None
        finally:
            _lib.load().helio_set_bwd_precision(bwd_mode)
            _lib.load().helio_set_fwd_precision(2)
        if snap_t is not None:
            try:
                tf32x3["default_formats_vs_3xtf32"] = _compare_snapshots(torch, _snapshot_step(torch, ENV[0], action0), snap_t)
            except Exception as e:
                tf32x3["default_formats_vs_3xtf32"] = dict(error=repr(e))
            del snap_t
    small = None
    if world == 1 and not args.no_small_field:
        try:
            small = small_field_numbers(torch, dev)
        except Exception as e:  # the headline line must still print
            small = dict(error=repr(e))

    line = dict(metric=METRIC, value=evals_step / (ms_step * 1e-3), unit=UNIT, n_gpus=world, steps=steps, warmup=warmup,
                ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                env_steps_per_s=1e3 / ms_step,
                config=dict(workload=WORKLOAD["name"], N=N, R=R, B_per_gpu=B, global_batch=B * world, parallelism=f"dp{world}",
                            splat=args.splat, target_cached=cached, check_finite="on (product default: asserts of step t raised at the next entry, no device sync)",
                            l2="inputs larger than L2 (>= 3 GB of images per step)",
                            renders_per_step="noisy fwd+bwd (target image cached: exact, it depends on the suns only)" if cached
                            else "noisy fwd+bwd, target fwd (re-rendered every step as in the reference)"),
                clocks=clocks,
                e2e=dict(value=evals_step / (ms_e2e * 1e-3), unit=UNIT, ms_per_step=ms_e2e,
                         h2d_bytes_per_step=h_action.numel() * 4, d2h_bytes_per_step=h_grad.numel() * 4 + 16,
                         path="caller-side copies around a device step" if args.e2e_serial else
                              "env.step(host action): sliced H2D under the forward, gradient D2H under the backward slices",
                         host_link=dict(h2d_gbs_min_over_ranks=h2d_gbs, d2h_gbs_min_over_ranks=d2h_gbs, concurrent_ranks=world,
                                        cpus_allowed=len(os.sched_getaffinity(0)),
                                        note="pinned 98 MB copies issued by all ranks at once, slowest rank")),
                gpu_launches=launches, roofline=roofline, cpu_baseline=cpu_baseline, gpu_eager_baseline=gpu_eager,
                small_field=small, culled=culled, both_3xtf32=tf32x3, strong_scaling=strong)
    line["uncached" if cached else "cached"] = other
    print(json.dumps(line), flush=True)
    if world > 1:
        _leave_process_group(torch, dist, [ENV[0]])
    return 0


if __name__ == "__main__":
    a = parse()
    sys.exit(main_reference(a) if a.impl == "reference" else main_ours(a))
