#!/usr/bin/env python
"""Stall-reason breakdown (warp state sampling) and a few utilisation metrics per kernel of an .ncu-rep.
    python tools/ncu_stalls.py report.ncu-rep"""
import csv
import io
import subprocess
import sys

txt = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[0]
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')][:70]
    print('==', name)
    stalls = []
    for i, h in enumerate(hdr):
        if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio'):
            try:
                stalls.append((float(r[i]), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
            except ValueError:
                pass
    tot = sum(v for v, _ in stalls)
    for v, n in sorted(stalls, reverse=True)[:9]:
        print('   %-28s %6.2f  (%4.1f%%)' % (n, v, 100 * v / max(tot, 1e-9)))
    for key in ('gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
                'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed_op_shared_ld.sum',
                'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum', 'smsp__inst_executed_op_global_red.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
                'launch__registers_per_thread', 'dram__bytes_read.sum', 'dram__bytes_write.sum'):
        if key in hdr:
            print('   %-70s %s' % (key, r[hdr.index(key)]))
