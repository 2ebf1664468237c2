#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
   python tools/launch_summary.py launches.csv ["# header line" ...]"""
import csv, re, sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, data = rows[0], rows[1:]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = {}
for r in data:
    name = re.sub(r"\(.*", "", r[ik])[:60]
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
    t = float(r[iv].replace(",", "")) * scale
    n, tot = agg.get(name, (0, 0.0))
    agg[name] = (n + 1, tot + t)
total = sum(t for _, t in agg.values())
for line in sys.argv[2:]:
    print(line)
for name, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{name:62s} n={n:4d} total={t:10.3f} ms share={t / total:6.1%} avg={t / n:8.3f} ms")
