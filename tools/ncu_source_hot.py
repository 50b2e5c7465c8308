#!/usr/bin/env python3
"""Hot source lines of an ncu capture:
   ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | python tools/ncu_source_hot.py [N]
SASS rows are attributed to the CUDA source line they follow; prints instructions executed (warp level),
stall samples and average active lanes per (file, line)."""
import csv
import sys

rows = list(csv.reader(sys.stdin))
cur_file, hdr, agg, cur_line = None, None, {}, None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or len(r) < 10:
        continue
    if r[0].strip().isdigit():
        cur_line = (int(r[0]), r[1].strip()[:110])
        continue
    if r[0].strip() != "" or cur_line is None:
        continue
    try:
        inst = float(r[hdr["Instructions Executed"]])
        tinst = float(r[hdr["Thread Instructions Executed"]])
        samp = float(r[hdr["# Samples"]])
    except (ValueError, KeyError):
        continue
    if inst == 0 and samp == 0:
        continue
    k = (cur_file, cur_line[0])
    a = agg.setdefault(k, [0.0, 0.0, 0.0, cur_line[1]])
    a[0] += inst
    a[1] += tinst
    a[2] += samp
tot = sum(a[0] for a in agg.values()) or 1
tots = sum(a[2] for a in agg.values()) or 1
print(f"total warp inst {tot/1e6:.1f} M, samples {tots:.0f}")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:n]:
    lanes = a[1] / a[0] if a[0] else 0
    print(f"{k[0]:20s}:{k[1]:4d} inst {100*a[0]/tot:5.1f}% samples {100*a[2]/tots:5.1f}% lanes {lanes:4.1f} | {a[3]}")
