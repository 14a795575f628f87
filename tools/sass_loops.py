#!/usr/bin/env python3
"""List the loops (backward branches) of a SASS dump with their body size and opcode histogram.
usage: cuobjdump -sass -fun <mangled> file.o | python tools/sass_loops.py [min_len]"""
import re
import sys
from collections import Counter

minlen = int(sys.argv[1]) if len(sys.argv) > 1 else 60
ins = []
for ln in sys.stdin:
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_index = {a: i for i, (a, _) in enumerate(ins)}
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA(?:\.\w+)*\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr_index:
            j = addr_index[tgt]
            body = ins[j:i + 1]
            if len(body) >= minlen:
                ops = Counter()
                for _, tt in body:
                    toks = [x for x in tt.split() if not x.startswith("@")]
                    ops[toks[0].split(".")[0]] += 1
                print(f"loop {tgt:#x}..{a:#x}: {len(body)} instrs  " + " ".join(f"{k}:{v}" for k, v in ops.most_common(14)))
