#!/usr/bin/env python3
"""Top stalled SASS instructions of an .ncu-rep (source page), as text."""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']]) for r in data)
print(rows[0][1][:100]); print('total samples', tot)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(r[ix[h]]) for r in data) for h in stalls}
print('by reason:', ', '.join(f"{h[6:]}={100*v/tot:.1f}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:topn]:
    s = int(r[ix['# Samples']])
    st = sorted(((int(r[ix[h]]), h[6:]) for h in stalls), reverse=True)[:2]
    print(f"{100*s/tot:5.1f}% {r[ix['Instructions Executed']]:>10} {r[ix['Source']][:72]:72s} {st}")
