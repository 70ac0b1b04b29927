#!/usr/bin/env python3
"""Static SASS instruction count per source line of one kernel (nvdisasm --print-line-info -c x.cubin).
usage: sass_lines.py x.sass <function substring> <source file substring> [lo hi]
Prints instructions attributed to each source line of that file and the total inside [lo, hi]."""
import re, sys, collections
path, fn, src = sys.argv[1:4]
lo = int(sys.argv[4]) if len(sys.argv) > 4 else 0
hi = int(sys.argv[5]) if len(sys.argv) > 5 else 10**9
infn = False
cur = None
counts = collections.Counter()
ops = collections.defaultdict(collections.Counter)
for l in open(path):
    if l.startswith("//---") and ".text." in l:
        infn = fn in l
        cur = None
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', l)
    if m and cur:
        counts[cur] += 1
        ops[cur][m.group(2).split(".")[0]] += 1
tot = 0
for (f, n), c in sorted(counts.items()):
    if src in f:
        inr = lo <= n <= hi
        tot += c if inr else 0
        print(f"{n:5d} {c:4d} {'*' if inr else ' '} " + " ".join(f"{k}:{v}" for k, v in ops[(f, n)].most_common()))
other = sum(c for (f, n), c in counts.items() if src not in f)
print("in range:", tot, " other files:", other, " total:", sum(counts.values()))
