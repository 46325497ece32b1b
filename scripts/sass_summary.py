#!/usr/bin/env python3
"""profiles/<tag>_sass_summary.txt: which Blackwell instructions the shipped library contains (cuobjdump -sass) and the
register / shared-memory / spill figures ptxas reports for the tcgen05 kernels (csrc/build.log).  Runs on the CPU box.

    python scripts/sass_summary.py r02
"""
import collections, os, re, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "doodle_b200", "libhelio_sm100.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
MN = ["UTCHMMA.2CTA", "UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "SYNCS.ARRIVE", "SYNCS.PHASECHK", "FENCE.VIEW.ASYNC",
      "MEMBAR.ALL.CTA", "MUFU.EX2", "MUFU.LG2", "STS.128", "STS.64", "LDS.128", "LDG.E.128", "STG.E.128", "HMMA", "ATOMG", "RED.E"]
per = collections.OrderedDict()
fn = None
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        per.setdefault(fn, collections.Counter())
        continue
    if fn is None:
        continue
    for k in MN:
        if re.search(r"\b" + re.escape(k) + r"\b", ln):
            per[fn][k] += 1
            break
out = [f"cuobjdump -sass doodle_b200/libhelio_sm100.so  (sm_100a; built by `make -C doodle_b200/csrc`)", ""]
tot = collections.Counter()
for c in per.values():
    tot.update(c)
out.append("whole library: " + ", ".join(f"{k} {tot[k]}" for k in MN if tot[k]))
out.append("(UTCHMMA = tcgen05.mma, .2CTA = cta_group::2; LDTM = tcgen05.ld; UTCBAR = tcgen05.commit; UTCATOMSWS = tcgen05.alloc;")
out.append(" no UTMALDG / UTMASTG: operands are generated in registers, not loaded by TMA; no HMMA: no mma.sync / wmma path)")
out.append("")
out.append(f"{'kernel':70s} " + " ".join(f"{k:>12s}" for k in MN[:8]))
for fn, c in per.items():
    if "tc_kernel" in fn:
        out.append(f"{fn[:70]:70s} " + " ".join(f"{c[k]:12d}" for k in MN[:8]))
out.append("")
out.append("ptxas -v (csrc/build.log): registers / spills of the tcgen05 kernels")
log = open(os.path.join(ROOT, "doodle_b200", "csrc", "build.log")).read().splitlines()
for i, ln in enumerate(log):
    m = re.search(r"Compiling entry function '(\S+)'", ln)
    if m and "tc_kernel" in m.group(1):
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        info = " ".join(x.strip() for x in log[i + 1:i + 4] if "spill" in x or "Used" in x)
        info = re.sub(r"ptxas info\s*:\s*", "", info)
        out.append(f"  {name[:64]:64s} {info}")
p = os.path.join(ROOT, "profiles", f"{tag}_sass_summary.txt")
open(p, "w").write("\n".join(out) + "\n")
print("\n".join(out[:14]))
