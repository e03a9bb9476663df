#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` per CUDA source
line: stall samples, executed warp instructions and the dominant stall reasons.
usage: ncu_lines.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys


def num(v):
    try:
        return int(v)
    except (TypeError, ValueError):
        return 0

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
cur_file = ''
lines = {}
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if r[0] == 'Line No':
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or r[0] == '':
        continue
    d = dict(zip(hdr[4:], r[4:]))
    key = (cur_file, int(r[0]))
    samples = num(d['# Samples'])
    inst = num(d['Instructions Executed'])
    stalls = {k[6:]: num(v) for k, v in d.items() if k.startswith('stall_') and 'Not Issued' not in k}
    e = lines.setdefault(key, dict(src=r[1], samples=0, inst=0, stalls={}))
    e['samples'] += samples
    e['inst'] += inst
    for k, v in stalls.items():
        e['stalls'][k] = e['stalls'].get(k, 0) + v
tot_s = sum(e['samples'] for e in lines.values()) or 1
tot_i = sum(e['inst'] for e in lines.values()) or 1
print('total samples %d, total warp instructions %d' % (tot_s, tot_i))
agg = {}
for e in lines.values():
    for k, v in e['stalls'].items():
        agg[k] = agg.get(k, 0) + v
print('stall mix: ' + ', '.join('%s %.1f%%' % (k, 100. * v / tot_s)
                               for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
print('%-22s %7s %7s  %-44s %s' % ('file:line', 'samp%', 'inst%', 'top stalls', 'source'))
for key, e in sorted(lines.items(), key=lambda kv: -kv[1]['samples'])[:top]:
    st = ', '.join('%s %d' % (k, v) for k, v in sorted(e['stalls'].items(), key=lambda kv: -kv[1])[:3] if v)
    print('%-22s %6.2f%% %6.2f%%  %-44s %s' % ('%s:%d' % key, 100. * e['samples'] / tot_s,
                                            100. * e['inst'] / tot_i, st[:44], e['src'].strip()[:90]))
