#!/usr/bin/env python3
"""DRAM traffic of the dominant kernel per cell crossing, from a named .ncu-rep (ncu --set full), written to
profiles/traffic.json — the `roofline.traffic` figure of bench.py (dram__bytes_read.sum + dram__bytes_write.sum of
ONE launch, divided by the crossings that launch made, so that bench.py can scale it to the crossings of its own
launches).

usage: ncu_traffic.py WORKLOAD REPORT.ncu-rep KERNEL_REGEX CROSSINGS_OF_THE_CAPTURED_LAUNCH [launch index]
The crossings come from the log of the same capture (tools/profile_shoot.py prints packets and crossings/packet)."""
import csv, json, re, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def main():
    workload, rep, kre, crossings = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    which = sys.argv[5] if len(sys.argv) > 5 else "0"   # launch index, or "all": sum over every captured launch
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    hits = [r for r in rows[2:] if re.search(kre, r[ik])]
    sel = hits if which == "all" else [hits[int(which)]]
    r = sel[0]

    def val(name):
        i = hdr.index(name)
        u = units[i].lower()
        return sum(float(x[i].replace(",", "")) for x in sel) * {"byte": 1., "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1.)
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    p = ROOT / "profiles" / "traffic.json"
    d = json.loads(p.read_text()) if p.exists() else {}
    d[workload] = {"kernel": r[ik].split("(")[0].replace("void ", ""), "report": Path(rep).name,
                   "launches_summed": len(sel), "dram_bytes_read": rd, "dram_bytes_write": wr, "crossings_of_the_launch": crossings,
                   "dram_bytes_per_crossing": (rd + wr) / crossings,
                   "duration_ms": val("gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "msecond": 1., "ms": 1., "nsecond": 1e-6}.get(units[hdr.index("gpu__time_duration.sum")], 1.)}
    p.write_text(json.dumps(d, indent=1, sort_keys=True) + "\n")
    print(workload, d[workload])


if __name__ == "__main__":
    main()
