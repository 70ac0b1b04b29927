#!/usr/bin/env python3
"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`) of
`bench.py --steps K`: per kernel the launches, total time and share inside the K timed iterations (the launches
between the last K+1 state updates, or the K iterations from state update number FIRST of the list on).
Usage: python tools/launch_summary.py launches.csv K [FIRST] > profiles/x.md"""
import csv
import sys
from collections import OrderedDict


def main():
    path, steps = sys.argv[1], int(sys.argv[2])
    first_update = int(sys.argv[3]) if len(sys.argv) > 3 else None   # index of the first timed state update in the list
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    launches = [(r[4].split("(")[0].replace("void ", "").split("<")[0], float(r[14].replace(",", "")) * 1e-6) for r in rows]
    upd = [i for i, (n, _) in enumerate(launches) if n.startswith("update_temperature") or n.startswith("update_state")]
    if first_update is None:
        first = upd[-steps - 1] + 1 if len(upd) > steps else 0
        timed = launches[first:upd[-1] + 1]
    else:
        first = upd[first_update - 1] + 1 if first_update > 0 else 0
        timed = launches[first:upd[first_update + steps - 1] + 1]
    agg = OrderedDict()
    for n, ms in timed:
        c, t = agg.get(n, (0, 0.))
        agg[n] = (c + 1, t + ms)
    total = sum(t for _, t in agg.values())
    print(f"{len(launches)} launches in the list, {len(timed)} inside the {steps} timed iterations.\n")
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {n} | {c} | {t:.2f} | {100 * t / total:.1f} % |")
    print(f"\nSum {total:.1f} ms for {steps} iterations.")


if __name__ == "__main__":
    main()
