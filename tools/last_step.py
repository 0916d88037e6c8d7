"""Prints the launches of the LAST of `steps` identical steps from an ncu launch-list CSV (tools/step_launches.py runs)."""
import csv
import sys

path, steps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 4
rows = [r for r in csv.reader(open(path)) if len(r) > 5]
h, rows = rows[0], rows[1:]
ki, vi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
n = len(rows) // steps
tot = 0.0
for r in rows[-n:]:
    us = float(r[vi].replace(",", "")) / 1e3
    tot += us
    print(f"{us:9.1f} us  {r[gi]:>16}  {r[ki].split('(')[0][:90]}")
print(f"{tot:9.1f} us  total over {n} launches")
