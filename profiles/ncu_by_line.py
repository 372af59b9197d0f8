#!/usr/bin/env python
"""Aggregate an ncu report's executed-instruction counts per CUDA source line.
usage: python profiles/ncu_by_line.py report.ncu-rep [top_n]"""
import csv, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
fname = None; hdr = None; cur = None
agg = defaultdict(lambda: [0, 0, 0, ""])   # (file,line) -> [warp instr, thread instr, stall samples, text]
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) == 2 and r[0] == "Function Name": continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 10: continue
    if r[0] != "":
        cur = (fname, int(r[0])); agg[cur][3] = r[1]
    elif cur is not None:
        try:
            agg[cur][0] += int(r[7]); agg[cur][1] += int(r[8]); agg[cur][2] += int(r[4])
        except ValueError:
            pass
tot = sum(v[0] for v in agg.values()); tots = sum(v[2] for v in agg.values())
print(f"total warp instructions {tot}, stall samples {tots}")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if v[0] == 0: break
    print(f"{100*v[0]/tot:5.1f}% instr {100*v[2]/max(tots,1):5.1f}% stall  thr={v[1]/max(v[0],1):4.1f}  {f}:{l:<4d} {v[3].strip()[:100]}")
