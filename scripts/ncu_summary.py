#!/usr/bin/env python
"""Summarise .ncu-rep captures (ncu --set full) into the JSON kept under profiles/.

    python scripts/ncu_summary.py OUT.json name=report.ncu-rep [name=report.ncu-rep ...]

Reads each report with `ncu -i ... --page raw --csv` (no GPU needed) and keeps the metrics DESIGN.md quotes:
duration, grid, registers, instruction count, issue / pipe utilisation, DRAM bytes, L2 hit rate and the top
warp-stall reasons per issued instruction.
"""
from __future__ import annotations

import csv
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"
STALL_SUFFIX = "_per_issue_active.ratio"


def summarise(report: str) -> dict:
    out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    rec = dict(zip(hdr, vals))
    unit = dict(zip(hdr, units))
    res = {"Kernel Name": rec.get("Kernel Name", "")}
    for k in KEEP:
        if k in rec and rec[k] not in ("", "n/a"):
            res[k] = {"unit": unit[k], "value": float(rec[k].replace(",", ""))}
    stalls = []
    for h in hdr:
        if h.startswith(STALL_PREFIX) and h.endswith(STALL_SUFFIX) and rec[h] not in ("", "n/a"):
            stalls.append((float(rec[h]), h[len(STALL_PREFIX):-len(STALL_SUFFIX)]))
    res["stall_per_issue"] = {name: round(v, 4) for v, name in sorted(stalls, reverse=True)[:6]}
    return res


def main(argv):
    if len(argv) < 3:
        print(__doc__)
        return 2
    out = {}
    for item in argv[2:]:
        name, _, path = item.partition("=")
        out[name] = summarise(path)
    with open(argv[1], "w") as fh:
        json.dump(out, fh, indent=1)
    for name, rec in out.items():
        t = rec.get("gpu__time_duration.sum", {})
        print(f"{name}: {rec['Kernel Name'][:60]}  {t.get('value')} {t.get('unit')}")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
