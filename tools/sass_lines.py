#!/usr/bin/env python
"""Static SASS histogram of one kernel by source line (no GPU needed): which lines of the .cuh files the instructions of a
kernel come from, plus an opcode histogram.  Straight-line per-pair code makes the static count a good proxy for the dynamic
per-pair cost that ncu's source counters report.

    python tools/sass_lines.py build/csrc/inst_dist_8.o 'render_kernelILi8ELi1ELb1' [--top 40] [--ops]
"""
import collections
import re
import subprocess
import sys
import tempfile
import os


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 40
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=d, check=True, capture_output=True)
        cubin = [os.path.join(d, f) for f in os.listdir(d) if f.endswith('.cubin')][0]
        txt = subprocess.run(['nvdisasm', '--print-line-info', cubin], check=True, capture_output=True, text=True).stdout
    cur_fn, cur_line, lines, ops = None, None, collections.Counter(), collections.Counter()
    n = 0
    for ln in txt.splitlines():
        m = re.match(r'\s*\.text\.(\S+):', ln)
        if m:
            cur_fn = m.group(1)
            continue
        if cur_fn is None or pat not in cur_fn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)', ln)
        if m:
            n += 1
            lines[cur_line] += 1
            ops[m.group(2).split('.')[0]] += 1
    print('kernel pattern %s: %d SASS instructions' % (pat, n))
    src = {}
    for (f, l), c in lines.most_common(top):
        if f not in src:
            p = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'gendr_b200', 'csrc', f)
            src[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = src[f][l - 1].strip()[:110] if l - 1 < len(src[f]) else ''
        print('%5d  %5.1f%%  %s:%d  %s' % (c, 100. * c / n, f, l, text))
    if '--ops' in sys.argv:
        print('opcodes:', ', '.join('%s %d' % kv for kv in ops.most_common(40)))


if __name__ == '__main__':
    main()
