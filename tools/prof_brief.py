#!/usr/bin/env python
"""Brief of one ncu capture directory (raw.csv + source.csv from tools/gpu_src.sh):
headline metrics, stall mix, top stalled SASS lines.  Usage: python tools/prof_brief.py <dir> [points] [ntop]"""
import csv, sys
d = sys.argv[1]; pts = float(sys.argv[2]) if len(sys.argv) > 2 else 1e8; ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 25
r = list(csv.reader(open(d + "/raw.csv"))); H, U, V = r[0], r[1], r[2]
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_sectors_srcunit_tex_op_red.sum", "lts__t_sectors_srcunit_tex_op_red_lookup_miss.sum", "lts__t_bytes.sum"]
for k in keys:
    if k in H: print(f"{k:80s} {V[H.index(k)]} {U[H.index(k)]}")
st = []
for i, h in enumerate(H):
    if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
        try: st.append((float(V[i]), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
        except ValueError: pass
print("stalls/issue:", ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:9]))
rows = list(csv.reader(open(d + "/source.csv"))); H = rows[1]
ai, si, ei, ss = H.index("Address"), H.index("Source"), H.index("Instructions Executed"), H.index("# Samples")
base = None; data = []
for x in rows[2:]:
    if len(x) <= ei: continue
    a = int(x[ai], 16) if x[ai].startswith("0x") else int(x[ai])
    if base is None: base = a
    data.append((a - base, x[si], int(x[ei] or 0), int(x[ss] or 0)))
te, ts = sum(x[2] for x in data), sum(x[3] for x in data)
print(f"{te/pts:.2f} warp-instr/pt; top stalled lines:")
for x in sorted(data, key=lambda x: -x[3])[:ntop]:
    print(f"{x[0]:6x} {100*x[3]/ts:5.2f}% exec/pt {x[2]/pts:6.3f}  {x[1][:90]}")
