#!/usr/bin/env python
"""Static instruction statistics of the main loop of a kernel's SASS (from cuobjdump -sass):
opcode mix and how many DFMAs read three distinct vector registers (3 FP64-pipe cycles instead of 2)."""
import re
import sys

lines = [l for l in open(sys.argv[1]) if re.search(r'/\*[0-9a-f]{4}\*/', l)]
best = None
for l in lines:
    m = re.search(r'/\*([0-9a-f]{4})\*/\s+(@!?U?P\d\s+)?BRA\S*\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)', l)
    if m:
        a, t = int(m.group(1), 16), int(m.group(3), 16)
        if t < a and (best is None or a - t > best[1] - best[0]):
            best = (t, a)
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
print("DFMA with 3 distinct vector-register sources:", three, " others:", le2)
print(sorted(byop.items(), key=lambda kv: -kv[1])[:16])
