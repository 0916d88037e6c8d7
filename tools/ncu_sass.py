#!/usr/bin/env python
"""Instruction-level view of one kernel in an ncu report: opcode histogram (by executed warp instructions) and the hottest SASS
lines by stall samples.   python tools/ncu_sass.py report.ncu-rep [top]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# several kernels may follow one another: split on the "Kernel Name" rows
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks:
    hdr, body = b["rows"][0], b["rows"][1:]
    ia, isrc, ismp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    print("==", b["name"][:100])
    ops = collections.Counter()
    total = 0
    for r in body:
        try:
            n = int(r[iex])
        except Exception:
            continue
        op = r[isrc].split()[0] if r[isrc].split() else "?"
        if op.startswith("@"):
            op = r[isrc].split()[1]
        ops[op.split(".")[0]] += n
        total += n
    print("total warp instructions", total)
    for op, n in ops.most_common(22):
        print(f"  {op:12s} {n:10d} {100.0 * n / total:5.1f}%")
    print("-- hottest lines by samples")
    body2 = sorted([r for r in body if r[ismp].isdigit()], key=lambda r: -int(r[ismp]))
    tot_s = sum(int(r[ismp]) for r in body2)
    for r in body2[:top]:
        print(f"  {int(r[ismp]):6d} {100.0 * int(r[ismp]) / max(tot_s, 1):5.1f}%  ex={r[iex]:>8s}  {r[isrc][:110]}")
