#!/usr/bin/env python
"""Top SASS lines of an `ncu --page source --csv` dump by stall samples: tools/ncu_hot_lines.py FILE [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = [i for i, r in enumerate(rows[:10]) if "Source" in r][0]
hdr = rows[hi]
cs, cn = hdr.index("Source"), hdr.index("# Samples")
stall = [j for j, h in enumerate(hdr) if h.startswith("stall_")]
data = []
for k, r in enumerate(rows[hi + 1:]):
    try:
        s = int(r[cn])
    except (ValueError, IndexError):
        continue
    top = sorted(((int(r[j] or 0), hdr[j]) for j in stall), reverse=True)[:2]
    data.append((s, k, r[cs][:90], top))
total = sum(d[0] for d in data)
print("total samples", total, "instructions", len(data))
acc = 0
for s, k, src, top in sorted(data, reverse=True)[:n]:
    print(f"{s:6d} {100.0 * s / total:5.1f}%  #{k:5d}  {src:90s} {top}")
