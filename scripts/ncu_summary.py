"""Summarise an `ncu --set full` report: one JSON line per kernel launch with the metrics the
roofline discussion uses.  Usage: python scripts/ncu_summary.py report.ncu-rep > summary.jsonl"""
import csv
import io
import json
import re
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "launch__shared_mem_per_block_static",
        "sm__cycles_active.min", "sm__cycles_active.avg", "sm__cycles_active.max", "gpc__cycles_elapsed.max"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
for r in data:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("vfmb::", "")
    out = {"kernel": name}
    for k in KEEP:
        if k in col:
            out[k] = r[col[k]] + " " + units[col[k]]
    print(json.dumps(out))
