#!/usr/bin/env python
"""Summarise the SASS source page of an .ncu-rep (ncu --set full --import-source on): warp-state samples by stall reason over the
whole kernel and the instructions that collect the most samples.   usage: ncu_source_summary.py report.ncu-rep [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# one section per profiled launch: a "Kernel Name" row, a header row, then one row per SASS instruction
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
seen = set()
for a, b in zip(starts, starts[1:]):
    key = (rows[a][1], b - a)
    if key in seen: continue          # ncu repeats a launch's section when several launches are in the report
    seen.add(key)
    print(rows[a][0] + ": " + rows[a][1])
    hdr, data = rows[a + 1], [r for r in rows[a + 2:b] if len(r) >= len(rows[a + 1])]
    ix = {h: i for i, h in enumerate(hdr)}
    num = lambda r, h: int(r[ix[h]] or 0)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    total = sum(num(r, "# Samples") for r in data)
    print(f"SASS instructions: {len(data)}, warp-state samples: {total}")
    agg = sorted(((sum(num(r, h) for r in data), h) for h in stalls), reverse=True)
    print("samples by warp state: " + ", ".join(f"{h[6:]} {100.0 * v / total:.1f}%" for v, h in agg if v))
    print(f"top {top_n} instructions by samples (address, samples, share, instruction, two largest states):")
    for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:top_n]:
        st = sorted(((num(r, h), h[6:]) for h in stalls), reverse=True)[:2]
        print(f"  {r[ix['Address']][-5:]}  {num(r, '# Samples'):6d}  {100.0 * num(r, '# Samples') / total:5.2f}%  {r[ix['Source']].strip():60.60s}  "
              + ", ".join(f"{h} {v}" for v, h in st))
