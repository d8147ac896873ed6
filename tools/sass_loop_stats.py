#!/usr/bin/env python
"""Static instruction statistics of the main loop of one kernel's SASS:
    sass_loop_stats.py <lib.so or .sass> <kernel-name substring>
opcode mix of the largest backward-branch loop and how many DFMAs read three distinct vector
registers (3 FP64-pipe cycles instead of 2)."""
import re
import subprocess
import sys

src, want = sys.argv[1], sys.argv[2]
text = open(src).read() if src.endswith(".sass") else subprocess.run(["cuobjdump", "-sass", src], capture_output=True, text=True).stdout
funcs, cur = {}, None
for l in text.splitlines():
    m = re.match(r"\s*Function : (\S+)", l)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur:
        funcs[cur].append(l)
name = next(n for n in funcs if want in n)
lines = [l for l in funcs[name] if re.search(r'/\*[0-9a-f]{4}\*/', l)]
best = None
for l in lines:
    m = re.search(r'/\*([0-9a-f]{4})\*/\s+(@!?U?P\d\s+)?BRA\S*\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)', l)
    if m:
        a, t = int(m.group(1), 16), int(m.group(3), 16)
        if t < a and (best is None or a - t > best[1] - best[0]):
            best = (t, a)
print(name[:80])
print("loop", hex(best[0]), hex(best[1]), (best[1] - best[0]) // 16 + 1, "instructions")
three, le2, byop = 0, 0, {}
for l in lines:
    m = re.search(r'/\*([0-9a-f]{4})\*/\s+(?:@!?U?P\d\s+)?(\S+)\s+(.*?);', l)
    if not m or not best[0] <= int(m.group(1), 16) <= best[1]:
        continue
    op, args = m.group(2), m.group(3)
    base = op if op.startswith("IMAD.") and op.split('.')[1] in ("WIDE", "HI") else op.split('.')[0]
    byop[base] = byop.get(base, 0) + 1
    if base == 'DFMA':
        regs = {s.strip().lstrip('-|').rstrip('|').replace('.reuse', '') for s in args.split(',')[1:]}
        regs = {r for r in regs if re.match(r'R\d+$', r)}
        if len(regs) >= 3:
            three += 1
        else:
            le2 += 1
fp64 = sum(v for k, v in byop.items() if k in ("DFMA", "DADD", "DMUL", "DSETP"))
print("FP64 instructions:", fp64, " DFMA with 3 distinct vector-register sources:", three, " other DFMA:", le2)
print(sorted(byop.items(), key=lambda kv: -kv[1])[:20])
