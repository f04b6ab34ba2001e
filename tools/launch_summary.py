#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the LAST `n` launches.
usage: tools/launch_summary.py file.csv [n_last] [min_us_to_list]"""
import collections
import csv
import io
import re
import sys

path = sys.argv[1]
n_last = int(sys.argv[2]) if len(sys.argv) > 2 else 0
min_us = float(sys.argv[3]) if len(sys.argv) > 3 else 1e9
lines = [ln for ln in open(path).read().splitlines() if ln.startswith('"')]
seq = []
for d in csv.DictReader(io.StringIO("\n".join(lines))):
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = re.sub(r"\(.*", "", re.sub(r"^void ", "", d["Kernel Name"]))
    seq.append((int(d["ID"]), k, d["Grid Size"], d["Block Size"], float(d["Metric Value"].replace(",", "")) / 1000))
last = seq[-n_last:] if n_last else seq
agg = collections.defaultdict(lambda: [0, 0.0])
for s in last:
    agg[s[1]][0] += 1
    agg[s[1]][1] += s[4]
print(f"{len(seq)} launches in the file, {len(last)} summarised, {sum(v[1] for v in agg.values()):.1f} us")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:9.1f} us {v[0]:4d}x {k}")
for s in last:
    if s[4] >= min_us:
        print(s)
