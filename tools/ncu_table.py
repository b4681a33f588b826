"""Markdown table of the headline metrics of every launch in .ncu-rep files (read with `ncu -i ... --page raw --csv`).
usage: ncu_table.py a.ncu-rep [b.ncu-rep ...]"""
import csv
import io
import re
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "µs", 1.0),
        ("dram__bytes_read.sum", "DRAM rd MB", 1.0),
        ("dram__bytes_write.sum", "DRAM wr MB", 1.0),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor %", 1.0),
        ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue %", 1.0),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1TEX %", 1.0),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %", 1.0),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1.0),
        ("launch__registers_per_thread", "regs", 1.0)]


def to_mb(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)


print("| kernel | " + " | ".join(c[1] for c in COLS) + " |\n|---|" + "---|" * len(COLS))
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")].replace("void ", "").replace("sky::", "")).replace("(int)", "")
        cells = []
        for key, _, _ in COLS:
            if key not in hdr:
                cells.append("-"); continue
            i = hdr.index(key); v = r[i]
            if "bytes" in key:
                cells.append("%.0f" % to_mb(v, units[i]))
            elif key.startswith("gpu__time"):
                f = float(v.replace(",", ""))
                f *= {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "second": 1e6}.get(units[i], 1.0)
                cells.append("%.0f" % f)
            else:
                cells.append(v.split(".")[0] if key.startswith("launch") else "%.1f" % float(v.replace(",", "")))
        print(f"| `{name}` | " + " | ".join(cells) + " |")
