#!/usr/bin/env python
"""Static spill report: LDL/STL instructions inside the loops of a kernel (no GPU needed).
    python tools/sass_spills.py <object.o> <kernel-name-substring> [min_loop_len]"""
import re
import subprocess
import sys
import tempfile
import os

obj, pat = sys.argv[1], sys.argv[2]
min_len = int(sys.argv[3]) if len(sys.argv) > 3 else 150
with tempfile.TemporaryDirectory() as d:
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=d, check=True, capture_output=True)
    cubin = [os.path.join(d, f) for f in os.listdir(d) if f.endswith('.cubin')][0]
    txt = subprocess.run(['nvdisasm', '-c', '--print-line-info', cubin], check=True, capture_output=True, text=True).stdout
for part in re.split(r'\n(?=\.text\.)', txt):
    m = re.match(r'\.text\.(\S+):', part)
    if not m or pat not in m.group(1):
        continue
    labels, instrs, cur = {}, [], None
    for ln in part.splitlines():
        mm = re.match(r'^(\.L_x_\d+):', ln)
        if mm:
            labels[mm.group(1)] = len(instrs); continue
        mm = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if mm:
            cur = (mm.group(1).split('/')[-1], int(mm.group(2))); continue
        mm = re.match(r'\s*/\*([0-9a-f]+)\*/\s+(.*?);', ln)
        if mm:
            instrs.append((mm.group(2), cur))
    nl = sum(1 for t, _ in instrs if 'LDL' in t); ns = sum(1 for t, _ in instrs if 'STL' in t)
    print('%s: %d instructions, LDL %d STL %d, CALL %d' % (m.group(1)[:90], len(instrs), nl, ns, sum(1 for t, _ in instrs if t.startswith('CALL') or ' CALL' in t)))
    for i, (t, l) in enumerate(instrs):
        mm = re.search(r'BRA\S*\s+.*?(\.L_x_\d+)', t)
        if mm and mm.group(1) in labels and labels[mm.group(1)] <= i and i - labels[mm.group(1)] >= min_len:
            j = labels[mm.group(1)]
            seg = instrs[j:i + 1]
            print('   loop [%d..%d] len %d back-edge at %s: LDL %d STL %d' % (j, i, i - j, l, sum(1 for t2, _ in seg if 'LDL' in t2), sum(1 for t2, _ in seg if 'STL' in t2)))
