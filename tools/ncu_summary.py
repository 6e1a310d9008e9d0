#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): key throughput/stall metrics per kernel.
Usage: python tools/ncu_summary.py <report.ncu-rep> [out.csv]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'lts__t_sectors_srcunit_tex_op_red.sum',
        'lts__t_sectors_srcunit_tex_op_red_lookup_miss.sum', 'l1tex__t_sectors.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed_op_global_red.sum', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__thread_inst_executed_per_inst_executed.ratio']
stalls = [h for h in hdr if h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('.ratio')] or \
         [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') ]
idx = {h: i for i, h in enumerate(hdr)}
out = []
for r in rows[2:]:
    print('-' * 100)
    for w in want:
        if w in idx:
            print(f"{w:75s} {r[idx[w]]:>20s} {units[idx[w]]}")
            out.append((r[idx['Kernel Name']][:60], w, r[idx[w]], units[idx[w]]))
    st = []
    for h in hdr:
        if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio'):
            try: st.append((float(r[idx[h]]), h))
            except ValueError: pass
    for v, h in sorted(st, reverse=True)[:8]:
        nm = h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')
        print(f"   stalled warps per issue: {nm:30s} {v:8.3f}")
        out.append((r[idx['Kernel Name']][:60], 'stall_' + nm, f"{v:.3f}", 'warps/issue'))
if len(sys.argv) > 2:
    with open(sys.argv[2], 'w', newline='') as fh:
        csv.writer(fh).writerows([('kernel', 'metric', 'value', 'unit')] + out)
