#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel share table (markdown)."""
import csv
import re
import sys
from collections import defaultdict

src, skip = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1 + skip:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").strip()
    t = float(r[vi].replace(",", ""))
    t = t / 1e3 if r[ui] == "ns" else t          # -> us
    agg[name][0] += 1
    agg[name][1] += t
tot = sum(v[1] for v in agg.values())
print(f"| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {t:.0f} | {100 * t / tot:.1f}% | {t / n:.1f} |")
print(f"| **total** | {sum(v[0] for v in agg.values())} | {tot:.0f} | 100% | |")
