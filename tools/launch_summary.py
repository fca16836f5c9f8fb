#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: launches, total ms, share, average us."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, start = r, i + 1
        break
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
d = collections.defaultdict(list)
for r in rows[start:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r[ui]]
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
    d[name].append(v)
tot = sum(sum(v) for v in d.values())
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
print("# kernel, launches, total ms, share %, avg us   (cold-cache serialised launches: compare shares)")
for k, v in sorted(d.items(), key=lambda x: -sum(x[1])):
    print(f"{k:64s} {len(v):5d} {sum(v):10.2f} {100 * sum(v) / tot:6.2f} {1e3 * sum(v) / len(v):10.1f}")
