#!/usr/bin/env python3
"""Condense an Nsight Compute report (.ncu-rep) into the per-kernel table kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/r01_xxx.md

Reads the report with `ncu -i <rep> --page raw --csv` (works without a GPU) and prints one markdown row per profiled
launch with the metrics the roofline discussion in DESIGN.md uses.
"""
import csv
import io
import subprocess
import sys

COLS = [
    ("Kernel Name", "kernel", None),
    ("Grid Size", "grid", None),
    ("Block Size", "block", None),
    ("gpu__time_duration.sum", "time_us", 1.0),
    ("dram__bytes_read.sum", "dram_rd_MB", 1.0),
    ("dram__bytes_write.sum", "dram_wr_MB", 1.0),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_%", 1.0),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%", 1.0),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_pipe_%", 1.0),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%", 1.0),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_%", 1.0),
    ("launch__registers_per_thread", "regs", 1.0),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem", 1.0),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts", 1.0),
    ("smsp__inst_executed.sum", "warp_insts", 1.0),
]


def main():
    rep = sys.argv[1]
    if rep.endswith(".csv"):      # already exported on the GPU box (`ncu -i rep --page raw --csv`): reports can exceed gpurun's 64 MiB
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = [(hdr.index(c), name) for c, name, _ in COLS if c in hdr]
    print("| " + " | ".join(f"{name} [{units[i]}]" if units[i] else name for i, name in idx) + " |")
    print("|" + "---|" * len(idx))
    for r in body:
        cells = []
        for i, name in idx:
            v = r[i]
            if name == "kernel":
                v = v.replace("<unnamed>::", "").replace("void ", "")[:70]
            else:
                try:
                    f = float(v.replace(",", ""))
                    v = f"{f:.0f}" if f >= 1000 else f"{f:.3g}"
                except ValueError:
                    pass
            cells.append(v)
        print("| " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
