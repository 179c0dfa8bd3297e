#!/usr/bin/env python
"""Turn the raw output of tools/gpu_r2_final.sh (gpurun_out/<tag>_*) into the tracked evidence under profiles/:
ncu summaries (JSON), stall breakdowns, source hot spots, launch list, bench lines, parity context numbers, sanitizer logs, the
roofline constants bench.py reads (profiles/roofline_traffic.json, keyed on a hash of the kernel sources) and a SASS opcode
histogram of the headline kernels.    python tools/collect_profiles.py <tag> <profile-prefix>"""
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
tag, prefix = sys.argv[1], sys.argv[2]
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')


def run(cmd):
    return subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT).stdout


def raw_rows(rep):
    rows = list(csv.reader(io.StringIO(run(['ncu', '-i', rep, '--page', 'raw', '--csv']))))
    hdr = rows[0]
    return [dict(zip(hdr, r)) for r in rows[2:]]


for name in ('bench_n1.json', 'bench_ref.json', 'bench_c4_n1.json', 'bench_c2_n1.json', 'parity.json', 'launches.csv', 'pytest_gpu.log',
             'small_configs_latency.json', 'sanitizer_memcheck.log', 'sanitizer_memcheck_ps.log', 'sanitizer_racecheck.log', 'smi.txt'):
    src = os.path.join(G, '%s_%s' % (tag, name))
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, '%s_%s' % (prefix, name)))
if os.path.exists(os.path.join(G, tag + '_parity.json')):
    shutil.copy(os.path.join(G, tag + '_parity.json'), os.path.join(P, 'parity_r2.json'))

import bench  # noqa: E402  (csrc_sha)
traffic = {'csrc_sha': bench.csrc_sha(), 'source': 'ncu --set full captures of tools/gpu_r2_final.sh (%s); per launch' % tag}
for wl, batch in (('c3', 64), ('c4', 2), ('c2', 16)):
    rep = os.path.join(G, '%s_%s_prof.ncu-rep' % (tag, wl))
    if not os.path.exists(rep):
        continue
    run([sys.executable, 'tools/ncu_summary.py', rep, os.path.join(P, '%s_%s_ncu_full_summary.json' % (prefix, wl))])
    open(os.path.join(P, '%s_%s_stalls.txt' % (prefix, wl)), 'w').write(run([sys.executable, 'tools/ncu_stalls.py', rep]))
    open(os.path.join(P, '%s_%s_source_hotspots.txt' % (prefix, wl)), 'w').write(run([sys.executable, 'tools/ncu_source_summary.py', rep, '40']))
    rows = raw_rows(rep)
    if len(rows) >= 2:
        fwd, bwd = rows[0], rows[1]

        def num(r, k):
            return float(r[k].replace(',', ''))

        def dram(r):      # ncu prints dram__bytes in a scaled unit: read the unit row via the summary tool instead
            return None
        summ = json.load(open(os.path.join(P, '%s_%s_ncu_full_summary.json' % (prefix, wl))))

        def to_bytes(txt):
            v, u = txt.split()[0], txt.split()[1] if len(txt.split()) > 1 else 'byte'
            mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
            return float(v) * mult
        e = {'batch': batch, 'forward_kernel': summ[0]['Kernel Name'], 'backward_kernel': summ[1]['Kernel Name']}
        for label, srow in (('forward', summ[0]), ('backward', summ[1])):
            e[label + '_dram_bytes_per_launch'] = to_bytes(srow['dram__bytes_read.sum']) + to_bytes(srow['dram__bytes_write.sum'])
            e[label + '_warp_instructions'] = float(srow['smsp__inst_executed.sum'].split()[0])
            e[label + '_active_lanes_per_instruction'] = float(srow['smsp__thread_inst_executed_per_inst_executed.ratio'].split()[0])
            e[label + '_issue_active_pct'] = float(srow['smsp__issue_active.avg.pct_of_peak_sustained_active'].split()[0])
            e[label + '_ms_under_ncu'] = srow['gpu__time_duration.sum']
        traffic[wl] = e
json.dump(traffic, open(os.path.join(P, 'roofline_traffic.json'), 'w'), indent=1)

# SASS opcode histogram of the headline kernels (static; proves the TMA bulk copy / mbarrier / red.global paths are in the binary)
out = []
for obj, pat, label in (('build/csrc/inst_dist_4.o', 'render_kernelILi4ELi3ELb0ELb1', 'C3 forward  render_kernel<gaussian, einstein, fwd, FAST>'),
                        ('build/csrc/inst_dist_4.o', 'render_kernelILi4ELi3ELb1ELb1', 'C3 backward render_kernel<gaussian, einstein, bwd, FAST> (pixel-stationary)'),
                        ('build/csrc/inst_dist_4.o', 'render_bwd_fs_kernelILi4ELi3ELb1', 'C3 backward render_bwd_fs_kernel<gaussian, einstein, FAST> (face-stationary)'),
                        ('build/csrc/inst_dist_8.o', 'render_kernelILi8ELi4ELb0ELb1', 'C4 forward  render_kernel<cauchy, yager2, fwd, FAST>'),
                        ('build/csrc/inst_dist_8.o', 'render_bwd_fs_kernelILi8ELi4ELb1', 'C4 backward render_bwd_fs_kernel<cauchy, yager2, FAST> (face-stationary)')):
    txt = run([sys.executable, 'tools/sass_lines.py', obj, pat, '--top', '0', '--ops'])
    ops = dict(re.findall(r'([A-Z0-9_]+) (\d+)', txt.split('opcodes:')[-1])) if 'opcodes:' in txt else {}
    full = run(['bash', '-c', "cuobjdump -sass %s | awk '/Function : .*%s/{f=1} f&&/Function : /&&!/%s/{f=0} f' | grep -oE '^\\s+/\\*[0-9a-f]+\\*/\\s+(@!?U?P[0-9T]+ )?[A-Z0-9_.]+' | awk '{print $NF}' | sort | uniq -c | sort -rn" % (obj, pat, pat)])
    keep = [ln for ln in full.splitlines() if re.search(r'UBLKCP|SYNCS|REDG|RED\.|MUFU|DFMA|DADD|DMUL|LDL|STL|SHFL|LDS|BAR|ATOM', ln)]
    out.append('== %s\n   %s\n   selected opcodes (count  mnemonic):\n%s\n' % (label, txt.splitlines()[0] if txt else '', '\n'.join('      ' + k.strip() for k in keep)))
open(os.path.join(P, '%s_sass_opcodes.txt' % prefix), 'w').write('\n'.join(out))
print('profiles written with prefix', prefix)
