#!/usr/bin/env python3
"""Instruction mix of the hot loops of a kernel, from `cuobjdump -sass` (no GPU needed).
usage: sass_loops.py file.o [substring of the function name]
For every backward branch whose body holds >= 32 DADD/DMUL it prints the body's opcode histogram
(bodies = address ranges [target, branch]); nested/overlapping bodies are all listed, largest first."""
import collections
import re
import subprocess
import sys

obj, pat = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
funcs, cur = {}, None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur:
        addr, ins = int(m.group(1), 16), m.group(2).strip()
        parts = ins.split()
        op = parts[1] if parts[0].startswith("@") else parts[0]
        funcs[cur].append((addr, op, ins))
for name, ins in funcs.items():
    if pat not in name:
        continue
    short = re.sub(r".*(jacobi\w+?)I(L.*?)EEv.*", r"\1<\2>", name).replace("Li", "").replace("E", ",")
    print("==", short, len(ins), "instructions")
    loops = []
    for a, op, full in ins:
        if op.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", full)
            if m and int(m.group(1), 16) <= a:
                loops.append((int(m.group(1), 16), a))
    for t, a in sorted(loops, key=lambda x: x[0] - x[1]):
        body = [(op, full) for (ad, op, full) in ins if t <= ad <= a]
        h = collections.Counter(op.split(".")[0] for op, _ in body)
        dp = h["DADD"] + h["DMUL"] + h["DFMA"]
        if dp < 32:
            continue
        lds = collections.Counter(op for op, _ in body if op.startswith("LDS") or op.startswith("STS"))
        print(f"  loop {t:#x}..{a:#x}: {len(body)} instr, DP {dp}, ", dict(h.most_common(14)), dict(lds))
