#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: per-opcode histogram is tools/sass_hist.py; this
prints executed instructions / stall samples for address ranges, and the top stalled instructions.
Usage: python tools/src_regions.py source.csv points [lo:hi:name ...]   (addresses hex, kernel-relative)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
pts = float(sys.argv[2])
H = rows[1]
ai, si, ei, ss = H.index("Address"), H.index("Source"), H.index("Instructions Executed"), H.index("# Samples")
data, base = [], None
for r in rows[2:]:
    if len(r) <= ei: continue
    a = int(r[ai], 16) if r[ai].startswith("0x") else int(r[ai])
    if base is None: base = a
    data.append((a - base, r[si], int(r[ei] or 0), int(r[ss] or 0)))
te, ts = sum(d[2] for d in data), sum(d[3] for d in data)
print(f"{len(data)} SASS lines, {te/pts:.2f} warp-instr/pt, {ts} samples")
for spec in sys.argv[3:]:
    lo, hi, name = spec.split(":")
    lo, hi = int(lo, 16), int(hi, 16)
    e = sum(d[2] for d in data if lo <= d[0] < hi); s = sum(d[3] for d in data if lo <= d[0] < hi)
    print(f"{name:30s} instr/pt {e/pts:7.2f} ({100*e/te:5.1f}%)  samples {100*s/ts:5.1f}%")
if len(sys.argv) == 3:
    # coarse automatic view: 64 equal address chunks
    n = len(data); step = max(1, n // 48)
    for i in range(0, n, step):
        ch = data[i:i + step]
        e = sum(d[2] for d in ch); s = sum(d[3] for d in ch)
        print(f"{ch[0][0]:6x}-{ch[-1][0]:6x} instr/pt {e/pts:6.2f} samples {100*s/ts:5.1f}%  {ch[0][1][:40]}")
