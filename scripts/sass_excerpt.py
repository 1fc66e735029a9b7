#!/usr/bin/env python3
"""Run here (no GPU): a short SASS excerpt of a kernel of librnmf_b200.so for profiles/ -- the producer's TMA issue loop
(UTMALDG + SYNCS.ARRIVE.TRANS64) and the first lines of the consumers' steady plane loop (SYNCS.PHASECHK.TRANS64.TRYWAIT, LDS,
DADD / DMUL, STS, STG).  usage: sass_excerpt.py <kernel name substring> [max lines]"""
import re
import subprocess
import sys

name = sys.argv[1]
max_lines = int(sys.argv[2]) if len(sys.argv) > 2 else 100
so = "rnmf_b200/librnmf_b200.so"
funcs = [l.split("Function :")[1].strip() for l in subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout.splitlines() if "Function :" in l]
fn = next(f for f in funcs if name in f)
out = subprocess.run(["cuobjdump", "-sass", "-fun", fn, so], capture_output=True, text=True).stdout.splitlines()
ins = [l for l in out if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", l)]
clean = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l).rstrip() for l in ins]
ops = [re.search(r"\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", l).group(2) if re.search(r"\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", l) else "" for l in clean]
hist = {}
for o in ops:
    k = o.split(".")[0]
    hist[k] = hist.get(k, 0) + 1
print(f"# {fn}")
print(f"# {len(ins)} SASS instructions; opcode counts of interest: " + ", ".join(f"{k} {hist.get(k, 0)}" for k in ("UTMALDG", "SYNCS", "LDS", "STS", "DADD", "DMUL", "DFMA", "STG", "LDG", "BRA")))
print("# full opcodes seen for TMA / mbarrier: " + ", ".join(sorted({o for o in ops if o.startswith("UTMALDG") or o.startswith("SYNCS")})))
# window 1: around the first UTMALDG
i = next(k for k, o in enumerate(ops) if o.startswith("UTMALDG"))
w1 = clean[max(0, i - 14): i + 14]
# window 2: the densest stretch of DADD/DMUL/LDS (the steady plane loop)
score = [1 if o.split(".")[0] in ("DADD", "DMUL", "LDS", "STS", "STG") else 0 for o in ops]
W = max_lines - len(w1) - 8
best, bs = 0, -1
run = sum(score[:W])
for k in range(0, len(score) - W):
    if run > bs:
        bs, best = run, k
    run += score[k + W] - score[k]
# start the window at the mbarrier wait that precedes it, if one is close
j = best
for k in range(best, max(0, best - 40), -1):
    if ops[k].startswith("SYNCS"):
        j = k
        break
print("\n# --- producer lane: TMA box loads of one plane, completion on the stage's mbarrier ---")
print("\n".join(w1))
print(f"\n# --- consumers: {W} instructions of a steady plane iteration (starting at the phase wait) ---")
print("\n".join(clean[j: j + W]))
