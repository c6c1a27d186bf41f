#!/usr/bin/env python
"""Per-source-line warp instructions, thread instructions (=> average active threads) and stall samples of one kernel.
usage: ncu_lines2.py source_page.csv [n_top]   (csv from: ncu -i rep --page source --csv --print-source cuda,sass)"""
import csv, collections, sys, os
rows = list(csv.reader(open(sys.argv[1]))); ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0, 0, 0]); tot = [0, 0, 0, 0]
for r in rows:
    if r and r[0] in ('File Path', 'File Name'): cur = r[1]; continue
    if r and r[0] == 'Line No':
        hdr = r; ci = hdr.index('Instructions Executed'); ti = hdr.index('Thread Instructions Executed'); si = hdr.index('# Samples'); li = hdr.index('stall_long_sb'); continue
    if not r or hdr is None or len(r) <= li or r[2] != '-': continue
    try: ln = int(r[0]); v = [int(r[ci]), int(r[ti]), int(r[si]), int(r[li])]
    except ValueError: continue
    a = agg[(cur, ln)]
    for k in range(4): a[k] += v[k]; tot[k] += v[k]
print('warp inst %d  thread inst %d  avg threads %.2f  samples %d (long_sb %.1f%%)' % (tot[0], tot[1], tot[1] / max(tot[0], 1), tot[2], 100.0 * tot[3] / max(tot[2], 1)))
src = {}
def line(f, ln):
    if f not in src:
        try: src[f] = open(f).read().split('\n')
        except OSError: src[f] = []
    return src[f][ln - 1].strip()[:90] if ln - 1 < len(src[f]) else ''
for (f, ln), (wi, thi, s, l) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:ntop]:
    print(f"{os.path.basename(f)[:14]:14s}:{ln:4d} winst {wi/tot[0]*100:5.2f}% thr/inst {thi/max(wi,1):5.1f} samp {s/max(tot[2],1)*100:5.2f}% lsb {l/max(tot[2],1)*100:5.2f}%  {line(f, ln)}")
