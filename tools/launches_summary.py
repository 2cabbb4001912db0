#!/usr/bin/env python
"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) by kernel and grid: count, mean / min / max device time."""
import collections, csv, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
c = collections.OrderedDict()
for row in csv.DictReader(lines):
    k = row["Kernel Name"].split("(")[0].replace("void ", "")
    c.setdefault((k, row["Grid Size"], row["Block Size"]), []).append(float(row["Metric Value"]))
print(f"{'launches':>8} {'avg us':>9} {'min us':>9} {'max us':>9}  kernel  grid x block")
for (k, g, b), v in c.items():
    print(f"{len(v):8d} {sum(v)/len(v)/1e3:9.1f} {min(v)/1e3:9.1f} {max(v)/1e3:9.1f}  {k}  {g} x {b}")
