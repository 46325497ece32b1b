#!/usr/bin/env python3
"""gpurun_out/scale_n{1,2,4,8}.log (bench.py lines of scripts/gpu_multi.sh) -> profiles/<tag>_scaling.{jsonl,md}.   python scripts/scaling_md.py r02"""
import glob, json, os, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lines = []
for f in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "scale_n*.log"))):
    for l in open(f):
        if l.startswith("{"):
            lines.append(json.loads(l))
lines.sort(key=lambda r: r["n_gpus"])
with open(os.path.join(ROOT, "profiles", f"{tag}_scaling.jsonl"), "w") as f:
    for r in lines:
        f.write(json.dumps(r) + "\n")
base = lines[0]
assert base["n_gpus"] == 1, "the table is relative to the N = 1 line"
out = ["# 1 -> 8 GPU scaling of the headline step (N=2000, R=256; `bench.py`, product defaults: target cached, f16x3 splats)", "",
       f"One 8xB200 box, `SCALE_NS=\"1 2 4 8\" SWEEP=1 bash scripts/gpu_multi.sh 8` ({base['steps']} timed steps per line, CUDA events, max over ranks).",
       "Weak = 4096 suns per GPU; strong = 4096 suns in total (`strong_scaling` block of the same line).  Efficiencies against the N = 1 line of the same box.",
       "The sharded env has no data-path collective: one all-reduce of the batch-mean metrics per step, and of 2 floats at reset.", "",
       "| GPUs | weak ms/step | weak evals/s | weak eff. | strong ms/step | strong evals/s | strong eff. | e2e ms/step | e2e evals/s | e2e eff. | H2D / D2H GB/s per GPU, all ranks at once |",
       "|---|---|---|---|---|---|---|---|---|---|---|"]
for r in lines:
    n = r["n_gpus"]; s = r.get("strong_scaling") or {}; e = r["e2e"]; hl = e.get("host_link") or {}
    st = (f"{s['ms_per_step']:.3f} | {s['value']:.4e} | {s['value'] / base['value']/ n:.3f}" if s.get("ms_per_step") else "— | — | —")
    out.append(f"| {n} | {r['ms_per_step']:.3f} | {r['value']:.4e} | {r['value'] / base['value'] / n:.3f} | {st} | {e['ms_per_step']:.3f} | {e['value']:.4e} | "
               f"{e['value'] / base['e2e']['value'] / n:.3f} | {hl.get('h2d_gbs_min_over_ranks', 0):.1f} / {hl.get('d2h_gbs_min_over_ranks', 0):.1f} |")
open(os.path.join(ROOT, "profiles", f"{tag}_scaling.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
