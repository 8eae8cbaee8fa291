#!/usr/bin/env python3
"""Prints the key counters of an .ncu-rep (first kernel): python tools_ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_sectors.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.per_cycle_active',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic']
for path in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for row in rows[2:3]:
        print('==', path, row[hdr.index('Kernel Name')][:60])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f'  {w:72s} {row[i]:>18s} {units[i]}')
        for i, h in enumerate(hdr):
            if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
                v = float(row[i].replace(',', ''))
                if v > 1.0:
                    print(f'    stall {h.split("issue_stalled_")[1].split("_per_issue")[0]:28s} {v:8.2f}')
