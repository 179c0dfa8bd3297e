"""Summarise `ncu --page source --csv --print-source cuda,sass` output: instructions executed and stall samples per
CUDA source line, per kernel.  Usage: python tools/ncu_source_summary.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
kernel, fname, hdr = None, None, None
agg = defaultdict(lambda: defaultdict(lambda: [0, 0, 0, '']))   # kernel -> (file,line) -> [inst, thread_inst, samples, src]
tot = defaultdict(lambda: [0, 0, 0])
cur_line = None
for row in csv.reader(io.StringIO(txt)):
    if not row:
        continue
    if row[0] in ('Kernel Name', 'Function Name'):
        kernel = row[1][:70]; continue
    if row[0] in ('File Name', 'File Path'):
        fname = row[1].split('/')[-1]; continue
    if row[0] == 'Line No':
        hdr = row; iI = hdr.index('Instructions Executed'); iT = hdr.index('Thread Instructions Executed'); iS = hdr.index('# Samples'); continue
    if hdr is None or kernel is None:
        continue
    if row[0].strip():              # a CUDA source line row (its own columns are the totals of the SASS rows below it)
        cur_line = (fname, int(row[0])); agg[kernel][cur_line][3] = row[1].strip()[:110]
        continue
    if len(row) > iT and row[2].startswith('0x'):
        try:
            ins, tins, smp = int(row[iI]), int(row[iT]), int(row[iS])
        except ValueError:
            continue
        a = agg[kernel][cur_line]; a[0] += ins; a[1] += tins; a[2] += smp
        t = tot[kernel]; t[0] += ins; t[1] += tins; t[2] += smp
for k in agg:
    t = tot[k]
    print('\n=== %s\n    warp-instr %.3e  thread-instr %.3e (avg %.1f active)  samples %d' % (k, t[0], t[1], t[1] / max(t[0], 1), t[2]))
    rows = sorted(agg[k].items(), key=lambda kv: -kv[1][0])[:top]
    for (f, ln), (ins, tins, smp, src) in rows:
        print('  %5.1f%% inst %5.1f%% smp  act %4.1f  %s:%d  %s' % (100.0 * ins / t[0], 100.0 * smp / max(t[2], 1), tins / max(ins, 1), f, ln, src))
