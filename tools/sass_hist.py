#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` dump by SASS opcode: executed warp instructions and
stall samples.  Usage: python tools/sass_hist.py dump.csv [points]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
pts = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
H = rows[1]
si, ei, ss = H.index("Source"), H.index("Instructions Executed"), H.index("# Samples")
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = samp = 0
for r in rows[2:]:
    if len(r) <= ei: continue
    op = r[si].split()
    if not op: continue
    o = op[1] if op[0].startswith("@") else op[0]
    o = ".".join(o.split(".")[:2]) if o.startswith(("LDS", "STS", "LDG", "STG", "RED", "ATOM")) else o.split(".")[0]
    e = int(r[ei] or 0); s = int(r[ss] or 0)
    agg[o][0] += e; agg[o][1] += s; agg[o][2] += 1
    tot += e; samp += s
print(f"total warp-instr {tot:.4g}  ({tot/pts:.2f} per point)  samples {samp}")
for o, (e, s, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"{o:14s} {e:14d} {100*e/tot:6.2f}%  per-pt {e/pts:7.3f}  samples {100*s/max(samp,1):6.2f}%  static {n}")
