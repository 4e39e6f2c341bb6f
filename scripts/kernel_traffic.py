"""profiles/kernel_traffic.json from `ncu --set full` reports: per captured kernel, the DRAM bytes one launch moved
(dram__bytes_read.sum + dram__bytes_write.sum), its duration, issue-slot utilisation and SM active / elapsed.
bench.py scales the per-cell figure of the capture nearest in size to the workload it runs.

usage: python scripts/kernel_traffic.py <tag> <report.ncu-rep | report.raw.csv>:<cells> [...]   (cells = grid cells the launch covered)
"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def to_bytes(v, unit):
    return float(v) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def to_us(v, unit):
    return float(v) * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6, "s": 1e6}[unit]


def main():
    tag = sys.argv[1]
    head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    caps = []
    for arg in sys.argv[2:]:
        path, cells = arg.rsplit(":", 1)
        cells = int(cells)
        # a report, or the `ncu -i <report> --page raw --csv` text the GPU session kept in its place
        raw = open(path).read() if path.endswith(".csv") else \
            subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        col = {k: i for i, k in enumerate(hdr)}
        for r in rows[2:]:
            def val(k):
                return r[col[k]].replace(",", "")
            rd = to_bytes(val("dram__bytes_read.sum"), units[col["dram__bytes_read.sum"]])
            wr = to_bytes(val("dram__bytes_write.sum"), units[col["dram__bytes_write.sum"]])
            dur = to_us(val("gpu__time_duration.sum"), units[col["gpu__time_duration.sum"]])
            name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
            name = name.split("::")[-1]
            caps.append({
                "kernel": name, "cells": cells, "source": f"profiles/{tag}_{Path(path).name.split('.')[0]}_ncu_details.txt (ncu --set full --clock-control none)",
                "dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes_per_cell_per_launch": (rd + wr) / cells,
                "duration_us_under_ncu": dur, "grid": val("launch__grid_size"), "block": val("launch__block_size"),
                "registers": val("launch__registers_per_thread"),
                "issue_active_pct": float(val("smsp__issue_active.avg.pct_of_peak_sustained_active")),
                "sm_active_over_elapsed": float(val("sm__cycles_active.avg")) / float(val("sm__cycles_elapsed.max")),
                "inst_executed": float(val("smsp__inst_executed.sum")),
                "warp_inst_per_cell": float(val("smsp__inst_executed.sum")) / cells,
            })
    out = {"build": head, "tag": tag, "captures": caps,
           "note": "one launch each; durations under ncu are cold-cache and serialised (bench.py measures the live ones)"}
    (ROOT / "profiles" / "kernel_traffic.json").write_text(json.dumps(out, indent=1) + "\n")
    for c in caps:
        print(c["kernel"][:40], c["cells"], f"{c['dram_bytes_per_cell_per_launch']:.2f} B/cell", f"{c['duration_us_under_ncu']:.1f} us",
              f"issue {c['issue_active_pct']:.1f}%", f"active/elapsed {c['sm_active_over_elapsed']:.3f}", f"{c['warp_inst_per_cell']:.2f} warp-inst/cell")


if __name__ == "__main__":
    main()
