#!/usr/bin/env python
"""Rank source lines of one kernel in an .ncu-rep by executed instructions and stall samples.
usage: ncu_lines.py report.ncu-rep [n_top] [regions] [kernel-regex]   (report captured with --import-source on)"""
import csv, collections, subprocess, sys, os, io
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cmd = ['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass']
if len(sys.argv) > 4: cmd += ['--kernel-name', 'regex:' + sys.argv[4]]
txt = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur = None; hdr = None; agg = collections.defaultdict(lambda: [0, 0, 0]); tot = [0, 0]
for r in rows:
    if r and r[0] == 'File Path': cur = r[1]; continue
    if r and r[0] == 'Function Name': continue
    if r and r[0] == 'Line No': hdr = r; li = hdr.index('stall_long_sb'); continue
    if not r or hdr is None or r[2] != '-': continue
    try: ln = int(r[0]); samples = int(r[4]); inst = int(r[7]); lsb = int(r[li])
    except ValueError: continue
    a = agg[(cur, ln)]; a[0] += samples; a[1] += inst; a[2] += lsb; tot[0] += samples; tot[1] += inst
print('total samples %d, warp instructions %d' % tuple(tot))
src = {}
def line(f, ln):
    if f not in src:
        try: src[f] = open(f).read().split('\n')
        except OSError: src[f] = []
    return src[f][ln - 1].strip()[:100] if ln - 1 < len(src[f]) else ''
for key in (1, 0):
    print('--- by', 'instructions' if key == 1 else 'stall samples')
    for (f, ln), (s, i, l) in sorted(agg.items(), key=lambda kv: -kv[1][key])[:ntop]:
        print(f"{os.path.basename(f)}:{ln:4d} inst {i/tot[1]*100:5.2f}% samp {s/tot[0]*100:5.2f}% (long_sb {l/tot[0]*100:5.2f}%)  {line(f, ln)}")
if len(sys.argv) > 3 and sys.argv[3]:   # region summary: "start:name,start:name,..." for the main .cu file
    regs = sorted((int(a.split(':')[0]), a.split(':')[1]) for a in sys.argv[3].split(','))
    out = collections.defaultdict(lambda: [0, 0])
    for (f, ln), (s, i, l) in agg.items():
        name = os.path.basename(f)
        if f.endswith('.cu'):
            for st, n in regs:
                if ln >= st: name = n
        out[name][0] += s; out[name][1] += i
    print('--- regions')
    for k, (s, i) in sorted(out.items(), key=lambda kv: -kv[1][1]): print(f"{k:28s} inst {i/tot[1]*100:5.1f}% samp {s/tot[0]*100:5.1f}%")
