#!/usr/bin/env python
"""Summarise an .ncu-rep (run here, no GPU needed): python tools_ncu_summary.py gpurun_out/prof.ncu-rep [kernel-regex]"""
import csv, subprocess, sys, re
rep = sys.argv[1]; pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size',
 'launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','smsp__inst_executed.sum',
 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
 'lts__t_sectors_srcunit_tex_op_read.sum','lts__t_sectors_srcunit_tex_op_write.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed',
 'lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    if pat and not pat.search(name): continue
    print("==", name[:110])
    for w in want:
        if w in hdr: print(f"  {w} = {r[hdr.index(w)]} {units[hdr.index(w)]}")
    st = []
    for i, h in enumerate(hdr):
        if 'average_warps_issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h:
            try: st.append((float(r[i]), h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')))
            except ValueError: pass
    print("  stalls (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
