#!/usr/bin/env python
"""stdin: `ncu --page source --csv --print-source sass` of one kernel.  Prints how many SASS instructions execute how
often (a histogram of per-instruction execution counts: the loop structure of the kernel at a glance) and the opcode mix
of the most executed group."""
import csv
import re
import sys
from collections import Counter

rows = list(csv.reader(sys.stdin))
hdr = next((r for r in rows if "Instructions Executed" in r), None)
if hdr is None:
    sys.exit("no SASS table on stdin")
ie = hdr.index("Instructions Executed")
data = [(r[1].strip(), int(r[ie])) for r in rows if len(r) > ie and r[ie].isdigit()]
tot = sum(n for _, n in data)
print(f"warp instructions executed: {tot}; SASS instructions: {len(data)}")
hist = Counter(n for _, n in data)
print("executions per instruction x instructions (share of all executed)")
for cnt, n in sorted(hist.items(), key=lambda x: -x[0] * x[1])[:12]:
    print(f"  {cnt:>12d} x {n:4d}  {100.0 * cnt * n / max(tot, 1):5.1f}%")


def opcode(s):
    s = re.sub(r"^@!?U?P\d+\s+", "", s)
    return s.split()[0].split(".")[0]


for cnt, n in sorted(hist.items(), key=lambda x: -x[0] * x[1])[:3]:
    mix = Counter(opcode(s) for s, c in data if c == cnt)
    print(f"opcodes of the {n} instructions executed {cnt} times: {dict(mix.most_common())}")
