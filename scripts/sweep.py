#!/usr/bin/env python3
"""BASELINE.json configs[4]: sweep N in {50,500,5000} x R in {64,128,512} x B in {25,1024,16384} (B per GPU, weak
scaling) of HelioEnv.step + backward, with the oracle port of the reference's CPU algorithm timed on a bounded sample
of each (N, R).  One JSON line per point on stdout (rank 0); run under torchrun for 2/4/8 GPUs.

    python scripts/sweep.py [--steps 3] [--cpu] [--points N,R,B ...] > gpurun_out/sweep_n1.jsonl
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--cpu", action="store_true", help="time the CPU oracle port on one sun per (N, R)")
    ap.add_argument("--points", nargs="*", default=None, help="N,R,B triples (default: the full 27-point grid)")
    ap.add_argument("--max-gb", type=float, default=150.0)
    ap.add_argument("--splat", default="auto", choices=["auto", "simt", "tc"])
    args = ap.parse_args()
    from doodle_b200 import HelioEnv, functional as Fn
    from doodle_b200.dist import make_sharded_env
    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.points:
        grid = [tuple(int(x) for x in p.split(",")) for p in args.points]
    else:
        grid = [(N, R, B) for N in (50, 500, 5000) for R in (64, 128, 512) for B in (25, 1024, 16384)]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tensor_peak = float(peaks.get("bf16_tflops_sustained") or 1400.0) / 3.0     # f16x3 (both splats by default): three fp16 MMAs per product, TFLOP/s
    hbm_peak = float(peaks.get("hbm_gbs") or 6500.0)
    cpu_cache = {}
    for (N, R, B) in grid:
        est_gb = (4 * B * R * R * 4 + 50 * B * N * 4 + 2.5 * B * R * R) / 1e9
        if est_gb > args.max_gb:
            if rank == 0:
                print(json.dumps(dict(N=N, R=R, B_per_gpu=B, skipped=f"needs ~{est_gb:.0f} GB")), flush=True)
            continue
        helio, targ_pos, targ_norm, area, _ = bench.make_inputs(N, B, rank=rank)
        torch.manual_seed(42 + rank)
        kw = dict(heliostat_pos=helio.to(dev), targ_pos=targ_pos.to(dev), targ_area=area, targ_norm=targ_norm.to(dev), sigma_scale=0.01,
                  error_scale_mrad=90.0, resolution=R, device=str(dev), new_errors_every_reset=True)   # product defaults: target cached,
        torch.cuda.synchronize(); t0 = time.perf_counter()                                             # finite check on, graph="auto"
        env = make_sharded_env(HelioEnv, global_batch_size=B * world, seed=42, **kw) if world > 1 else HelioEnv(batch_size=B, **kw)
        impl = dict(auto=0, simt=1, tc=2)[args.splat]
        env.noisy_field.splat_impl = env.ref_field.splat_impl = impl
        env.reset()
        torch.cuda.synchronize(); setup_s = time.perf_counter() - t0
        action0 = env.noisy_field.initial_action.detach().clone().view(B, N, 3)

        def step():
            a = action0.detach().requires_grad_(True)
            obs, m, mon = env.step(a)
            (m["mse"] + m["dist"] + m["bound"] + m["alignment_loss"]).backward()

        small = B * N * R * R < 2e12
        warm, steps = (5, 30) if small else (3, args.steps)
        for _ in range(warm):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms)
        replayed = getattr(env, "_step_graph", None) is not None
        Fn.reset_profile(True)                                   # per-kernel times: eager steps (event pairs cannot go into a graph)
        for _ in range(3):
            step()
        prof = Fn.collect_profile()
        Fn.reset_profile(False)
        evals = float(B) * world * N * R * R
        k = {n: round(v["avg_ms"] * 1e3, 1) for n, v in sorted(prof.items())}
        t_tensor = 6.0 * B * N * R * R / (tensor_peak * 1e12) * 1e3                  # noisy render fwd (2) + bwd (4) FLOP/eval; target cached
        t_hbm = (32.0 * B * R * R + 120.0 * B * N) / (hbm_peak * 1e9) * 1e3          # images: 1 write + 7 reads/writes of 4 B; per-(b,n) streams
        line = dict(N=N, R=R, B_per_gpu=B, n_gpus=world, splat=args.splat, ms_per_step=round(ms, 4), env_steps_per_s=round(1e3 / ms, 2),
                    evals_per_s=evals / (ms * 1e-3), roofline_ms=dict(tensor=round(t_tensor, 4), hbm=round(t_hbm, 4)),
                    frac_of_roofline=round(max(t_tensor, t_hbm) / ms, 4), bound="tensor" if t_tensor > t_hbm else "hbm",
                    tensor_model="f16x3: bf16_sustained/3", graph_replay=replayed, kernels_us=k, setup_s=round(setup_s, 3), mem_gb=round(torch.cuda.max_memory_allocated() / 1e9, 2))
        if args.cpu and rank == 0:
            if (N, R) not in cpu_cache:
                f = bench.cpu_step_factory(N, R, 1, os.cpu_count() or 1)
                t0 = time.perf_counter(); f(); dt = time.perf_counter() - t0
                cpu_cache[(N, R)] = (N * R * R / dt, dt)
            line["cpu_port_evals_per_s"] = cpu_cache[(N, R)][0]
            line["cpu_port_sample"] = f"1 sun, {cpu_cache[(N, R)][1]:.2f} s/step, {os.cpu_count()} threads"
            line["speedup_vs_cpu_port"] = round(line["evals_per_s"] / cpu_cache[(N, R)][0], 1)
        if rank == 0:
            print(json.dumps(line), flush=True)
        env.close()                                   # release captured graphs (they may hold NCCL kernels) before the next point
        del env, action0
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
    if world > 1:
        import threading
        torch.cuda.synchronize()
        t = threading.Timer(30.0, lambda: os._exit(0)); t.daemon = True; t.start()      # never let a teardown hang cost GPU time
        dist.destroy_process_group()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
