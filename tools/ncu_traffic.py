#!/usr/bin/env python
"""Write profiles/traffic.json from `ncu --set full` captures: DRAM bytes (read + write) per launch of a kernel.
    ncu_traffic.py <kernel>:<workload>=<file.ncu-rep> ..."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
out_path = os.path.join(ROOT, "profiles", "traffic.json")
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
for arg in sys.argv[1:]:
    key, rep = arg.split("=", 1)
    rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    rd = float(d["dram__bytes_read.sum"][1].replace(",", "")) * UNIT[d["dram__bytes_read.sum"][0]]
    wr = float(d["dram__bytes_write.sum"][1].replace(",", "")) * UNIT[d["dram__bytes_write.sum"][0]]
    out[key] = {"dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
                "kernel": d["Kernel Name"][1], "grid": d["launch__grid_size"][1], "report": os.path.basename(rep),
                "duration_ms_under_ncu": float(d["gpu__time_duration.sum"][1].replace(",", "")) *
                {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[d["gpu__time_duration.sum"][0]]}
json.dump(out, open(out_path, "w"), indent=1)
print(json.dumps(out, indent=1))
