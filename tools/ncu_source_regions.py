#!/usr/bin/env python3
"""Per-instruction view of one kernel of an .ncu-rep (source page, SASS): cumulative warp instructions, average
active threads and stall samples, printed per SASS instruction with running totals so that regions of the hot
loop can be read off.  usage: ncu_source_regions.py x.ncu-rep [min_share_percent]"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iS, iI, iT, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iI]) for r in rows[2:] if len(r) > iI and r[iI].isdigit())
tots = sum(int(r[iSm]) for r in rows[2:] if len(r) > iSm and r[iSm].isdigit())
print(f"total warp instructions {tot:.4g}, samples {tots}")
cum = 0
for k, r in enumerate(rows[2:]):
    if len(r) <= iI or not r[iI].isdigit():
        continue
    n, t, sm = int(r[iI]), int(r[iT]), int(r[iSm])
    cum += n
    top = sorted(((int(r[i]), h) for i, h in stall_cols if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:2]
    print(f"{k:5d} {100*n/tot:6.3f}% cum {100*cum/tot:6.2f}% thr {t/max(n,1):5.1f} smp {100*sm/max(tots,1):5.2f}%  {r[iS].strip()[:70]:70s} " + " ".join(f"{h[6:]}:{v}" for v, h in top))
