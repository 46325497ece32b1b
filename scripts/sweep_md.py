#!/usr/bin/env python3
"""jsonl of scripts/sweep.py -> markdown table under profiles/.   python scripts/sweep_md.py gpurun_out/sweep_n1.jsonl profiles/r02_sweep_n1.md"""
import json, os, sys
src, dst = sys.argv[1], sys.argv[2]
rows = [json.loads(l) for l in open(src) if l.startswith("{")]
g = rows[0].get("n_gpus", 1) if rows else 1
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    peaks = {}
TENSOR = float(peaks.get("bf16_tflops_sustained") or 1400.0) / 3.0       # f16x3: three fp16 MMAs per product
HBM = float(peaks.get("hbm_gbs") or 6500.0)


def roofline(r):
    """(bound, fraction) recomputed from N, R, B so that older jsonl files read against the same model."""
    N, R, B = r["N"], r["R"], r["B_per_gpu"]
    t_tensor = 6.0 * B * N * R * R / (TENSOR * 1e12) * 1e3
    t_hbm = (32.0 * B * R * R + 120.0 * B * N) / (HBM * 1e9) * 1e3
    return ("tensor" if t_tensor > t_hbm else "hbm"), max(t_tensor, t_hbm) / r["ms_per_step"]
out = [f"# Sweep (BASELINE.json configs[4]) on {g}x B200 -- `scripts/sweep.py` (B per GPU, weak scaling)", "",
       "One step = `HelioEnv.step` with the product defaults (target image cached, finite check on, `graph=\"auto\"`) + the caller's "
       "backward to `action.grad`; CUDA events, max over ranks.  `roofline` = max(tensor, HBM) time of the step: tensor = 6 FLOP/eval "
       "(noisy render forward + backward) at bf16_sustained/3 from MEASURED_PEAKS.json (both splats run f16x3 by default: three fp16 MMAs per fp32-accurate product); HBM = 32 B/pixel + "
       "120 B/(sun, heliostat) at the measured copy bandwidth.  `replay` = the step ran as CUDA-graph replays inside `env.step` "
       "(small fields).  Kernel times are from eager steps with the library's CUDA-event profiler.", "",
       "| N | R | B/GPU | ms/step | env-steps/s | evals/s (all GPUs) | bound | frac of roofline | replay | splat fwd us | splat bwd us | x cpu port |",
       "|---|---|---|---|---|---|---|---|---|---|---|---|"]
for r in rows:
    if "skipped" in r:
        out.append(f"| {r['N']} | {r['R']} | {r['B_per_gpu']} | skipped: {r['skipped']} | | | | | | | | |")
        continue
    k = r.get("kernels_us", {})
    bound, frac = roofline(r)
    out.append(f"| {r['N']} | {r['R']} | {r['B_per_gpu']} | {r['ms_per_step']:.3f} | {r['env_steps_per_s']:.1f} | {r['evals_per_s']:.3e} | {bound} | "
               f"{frac:.3f} | {'yes' if r.get('graph_replay') else 'no'} | {k.get('splat_fwd', '')} | {k.get('splat_bwd', '')} | "
               f"{r.get('speedup_vs_cpu_port', '')} |")
open(dst, "w").write("\n".join(out) + "\n")
print(f"{len(rows)} rows -> {dst}")
