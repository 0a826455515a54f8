#!/usr/bin/env python
"""Instruction / stall-sample share per source region of a kernel, from an ncu report with source.
   python tools/src_regions.py rep kernel-substr file.cu:lo-hi=name [...]   (inlined lines of other files: file:lo-hi too)"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
regions = []
for a in sys.argv[3:]:
    loc, name = a.split('=')
    f, rng = loc.split(':')
    lo, hi = rng.split('-')
    regions.append((f, int(lo), int(hi), name))
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = cur_fn = hdr = None
agg = {}
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': cur_fn = r[1]; continue
    if r[0] == 'Line No':
        hdr = r; iI = hdr.index('Instructions Executed'); iN = hdr.index('# Samples'); continue
    if hdr is None or cur_fn is None or kern not in cur_fn: continue
    if r[0].isdigit():
        try:
            ln = int(r[0]); ins = int(r[iI] or 0); smp = int(r[iN] or 0)
        except (ValueError, IndexError):
            continue
        name = 'other:' + cur_file
        for f, lo, hi, nm in regions:
            if f == cur_file and lo <= ln <= hi: name = nm; break
        a = agg.setdefault(name, [0, 0]); a[0] += ins; a[1] += smp
ti = sum(a[0] for a in agg.values()); ts = max(1, sum(a[1] for a in agg.values()))
print('total instr', ti, 'samples', ts)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print('%5.1f%% instr %5.1f%% samp  %s' % (100. * a[0] / ti, 100. * a[1] / ts, k))
