#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of metrics DESIGN.md cites.
usage: python profiles/ncu_summary.py gpurun_out/x.ncu-rep"""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__cycles_active.avg', 'sm__cycles_elapsed.max']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
H, U = rows[0], rows[1]
for r in rows[2:]:
    print('---', r[H.index('Kernel Name')][:90])
    for w in WANT:
        if w in H:
            print(f'   {w}: {r[H.index(w)]} {U[H.index(w)]}')
    st = []
    for i, h in enumerate(H):
        if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v > 0.08:
                st.append((v, h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
    print('   stalls/issue:', ', '.join(f'{n}={v:.2f}' for v, n in sorted(st, reverse=True)))
