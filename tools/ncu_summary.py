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
    w.writerow([""] + [("Mbyte" if units[idx[c]].endswith("byte") else units[idx[c]]) if c in idx else "" for c in COLS])
    # ncu picks the unit of a column from its largest value (byte .. Gbyte): bytes are normalised to Mbyte here
    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    for r in rows[2:]:
        vals = []
        for c in COLS:
            v = r[idx[c]] if c in idx else ""
            if c in idx and units[idx[c]] in scale and v:
                v = "%.6f" % (float(v.replace(",", "")) * scale[units[idx[c]]])
            vals.append(v)
        w.writerow([r[idx["Kernel Name"]]] + vals)


if __name__ == "__main__":
    main()
