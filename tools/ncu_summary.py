#!/usr/bin/env python3
"""ncu report -> one CSV line per launch with the metrics DESIGN.md / profiles/README.md quote.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.csv        (runs `ncu -i ... --page raw --csv` here, no GPU needed)
"""
import csv
import subprocess
import sys

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__grid_size"]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    w = csv.writer(sys.stdout)
    w.writerow(["Kernel Name"] + COLS)
    w.writerow([""] + [units[idx[c]] if c in idx else "" for c in COLS])
    for r in rows[2:]:
        w.writerow([r[idx["Kernel Name"]]] + [r[idx[c]] if c in idx else "" for c in COLS])


if __name__ == "__main__":
    main()
