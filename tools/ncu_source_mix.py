#!/usr/bin/env python
"""Per-opcode view of an `ncu --set full --import-source on` capture (source page, SASS):
    ncu_source_mix.py <file.ncu-rep>
For each opcode class: static count, warp-level instructions executed, share of all executed instructions, warp stall
samples that landed on it and the dominant stall reasons -- where the time of the kernel goes, by instruction kind."""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
print(rows[0][0], ":", rows[0][1][:110])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]


def klass(sass: str) -> str:
    ops = sass.split()
    op = ops[1] if ops and ops[0].startswith("@") else (ops[0] if ops else "?")
    if op.startswith("IMAD.WIDE") or op.startswith("IMAD.HI"):
        return ".".join(op.split(".")[:2])
    if op.startswith("IMAD.MOV") or op.startswith("IMAD.IADD") or op.startswith("IMAD.SHL"):
        return ".".join(op.split(".")[:2])
    return op.split(".")[0]


agg = defaultdict(lambda: [0, 0, 0, defaultdict(int)])
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    k = klass(r[ix["Source"]])
    a = agg[k]
    a[0] += 1
    a[1] += int(r[ix["Instructions Executed"]] or 0)
    a[2] += int(r[ix["# Samples"]] or 0)
    for s in stalls:
        a[3][s] += int(r[ix[s]] or 0)
tot_i = sum(a[1] for a in agg.values())
tot_s = sum(a[2] for a in agg.values())
print(f"{'opcode':<16}{'static':>7}{'warp insts':>14}{'% insts':>9}{'samples':>9}{'% samp':>8}  top stall reasons")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if a[1] == 0:
        continue
    top = sorted(a[3].items(), key=lambda kv: -kv[1])[:3]
    print(f"{k:<16}{a[0]:>7}{a[1]:>14}{100 * a[1] / tot_i:>9.2f}{a[2]:>9}{100 * a[2] / max(tot_s, 1):>8.2f}  "
          + ", ".join(f"{s[6:]} {n}" for s, n in top if n))
print(f"{'total':<16}{sum(a[0] for a in agg.values()):>7}{tot_i:>14}{100.0:>9.2f}{tot_s:>9}")
