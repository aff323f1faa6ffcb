#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): python scripts/ncu_summary.py gpurun_out/prof_X.ncu-rep"""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum', 'lts__t_bytes.sum']
STALL = 'smsp__average_warps_issue_stalled_'
rep = sys.argv[1]
rows = list(csv.reader(subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('== kernel', r[hdr.index('Kernel Name')][:60])
    for k in KEYS:
        if k in hdr:
            print('  %-75s %s %s' % (k, r[hdr.index(k)], units[hdr.index(k)]))
    st = [(float(r[i]), h) for i, h in enumerate(hdr) if h.startswith('smsp__average_warp') and h.endswith('_per_issue_active.ratio') and r[i]]
    for v, h in sorted(st, reverse=True)[:8]:
        print('  stall %-68s %.2f' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__average_warp_latency_issue_stalled_','').replace('_per_issue_active.ratio', ''), v))
