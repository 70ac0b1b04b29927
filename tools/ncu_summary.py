#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, without a GPU) into markdown: per captured launch the metrics the
roofline discussion needs.  Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [title] > profiles/x.md"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp instr"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe busy %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1TEX throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 (lts) throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__t_sectors.sum", "L2 sectors"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "RED lane-ops (l1tex red sectors)"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
]


def main():
    rep = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else rep
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    print(f"# {title}\n\nSource: `{rep}` (ncu --set full --clock-control none), read with `ncu -i … --page raw --csv`.\n")
    names = [r[ik].split("(")[0].replace("void ", "") for r in rows[2:]]
    print("| metric | " + " | ".join(f"{i}: {n}" for i, n in enumerate(names)) + " |")
    print("|---|" + "---|" * len(names))
    for key, label in KEYS:
        if key not in hdr:
            continue
        i = hdr.index(key)
        vals = []
        for r in rows[2:]:
            v = r[i]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.4g}"
            except ValueError:
                pass
            vals.append(f"{v} {units[i]}".strip())
        print(f"| {label} (`{key}`) | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
