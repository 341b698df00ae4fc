#!/usr/bin/env python3
"""Static view of the loops of one kernel in a cuobjdump -sass dump: for every backward branch, the size of the
loop body and its opcode histogram.   usage: sass_loops.py SASS.txt KERNEL_SUBSTRING [min_len]"""
import collections, re, sys

txt = open(sys.argv[1]).read()
want = sys.argv[2]
min_len = int(sys.argv[3]) if len(sys.argv) > 3 else 150
for part in re.split(r"\n\s+Function : ", txt)[1:]:
    name = part.split("\n", 1)[0]
    if want not in name:
        continue
    ins = [(int(a, 16), op, rest) for a, op, rest in
           re.findall(r"/\*([0-9a-f]{4,6})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)([^;]*);", part)]
    print(name[:90], len(ins), "instructions")
    for a, op, rest in ins:
        if op.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", rest)
            if m and int(m.group(1), 16) < a:
                t = int(m.group(1), 16)
                body = [o.split(".")[0] for (x, o, r) in ins if t <= x <= a]
                if len(body) >= min_len:
                    c = collections.Counter(body)
                    print(f"  loop {t:#x}..{a:#x}: {len(body)} instr  ", dict(c.most_common(14)))
