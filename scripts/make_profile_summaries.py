#!/usr/bin/env python3
"""Turn gpurun_out/{launches.csv, prof_*_tc.ncu-rep} into the tracked text summaries under profiles/.

    python scripts/make_profile_summaries.py r01
"""
import collections, csv, os, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(PR, exist_ok=True)

# ---- launch list -------------------------------------------------------------------------------
src = os.path.join(GO, "launches.csv")
if os.path.exists(src):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
    with open(os.path.join(PR, f"{tag}_bench_launches.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-small-field --no-culled --no-gpu-eager --no-both-3xtf32\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES with bench.py's live CUDA-event numbers\n")
        f.write("id,kernel,grid,block,duration_ns\n")
        for r in rows:
            f.write(f"{r[0]},\"{r[4].split('(')[0]}\",\"{r[8]}\",\"{r[7]}\",{r[-1]}\n")
    agg = collections.OrderedDict()
    for r in rows:
        a = agg.setdefault(r[4].split('(')[0][:80], [0, 0.0]); a[0] += 1; a[1] += float(r[-1])
    ours = {k: v for k, v in agg.items() if "helio::" in k}
    tot_ours = sum(v[1] for v in ours.values()); tot = sum(v[1] for v in agg.values())
    with open(os.path.join(PR, f"{tag}_bench_launches_summary.txt"), "w") as f:
        f.write(f"ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-small-field --no-culled --no-gpu-eager --no-both-3xtf32` (first 400 launches: env setup + warm-up + timed steps)\n")
        f.write(f"all kernels {tot/1e6:.2f} ms, libhelio kernels {tot_ours/1e6:.2f} ms ({100*tot_ours/tot:.1f} %)\n\n")
        f.write(f"{'launches':>8} {'total ms':>10} {'avg ms':>9} {'share':>7}  kernel\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:20]:
            f.write(f"{n:8d} {t/1e6:10.3f} {t/1e6/n:9.3f} {100*t/tot:6.1f}%  {k}\n")
    print(open(os.path.join(PR, f"{tag}_bench_launches_summary.txt")).read())

# ---- full captures -----------------------------------------------------------------------------
for what in ("fwd", "bwd"):
    rep = os.path.join(GO, f"prof_{what}_tc.ncu-rep")
    if not os.path.exists(rep):
        continue
    a = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    b = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_stalls.py"), rep, "25"], capture_output=True, text=True).stdout
    with open(os.path.join(PR, f"{tag}_splat_{what}_tc_ncu.txt"), "w") as f:
        f.write(f"ncu --set full --clock-control none --import-source on -k regex:splat_{what}_tc -s 3 -c 1 python scripts/prof_step.py\n")
        f.write("(N=2000, R=256, B=4096, product defaults: the kernel exactly as HelioEnv.step launches it in bench.py; one launch)\n\n== raw-page metrics ==\n" + a + "\n== source page: most-sampled SASS instructions ==\n" + b)
    print(what, "summary written")
