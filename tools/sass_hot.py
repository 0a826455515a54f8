#!/usr/bin/env python
"""Per-SASS-instruction hot spots of an ncu report (source page): python tools/sass_hot.py rep [kernel-substr] [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
blocks = out.split('"Kernel Name",')
for blk in blocks[1:]:
    lines = blk.split('\n')
    name = lines[0]
    if len(sys.argv) > 2 and sys.argv[2] not in name:
        continue
    rows = list(csv.reader(io.StringIO('\n'.join(lines[1:]))))
    h = rows[0]
    iA, iS, iN, iI, iT = h.index('Address'), h.index('Source'), h.index('# Samples'), h.index('Instructions Executed'), h.index('Avg. Threads Executed')
    data = []
    for r in rows[1:]:
        if len(r) < len(h):
            continue
        try:
            data.append((int(r[iI] or 0), int(r[iN] or 0), r[iT], r[iS], r[iA]))
        except ValueError:
            pass
    tot_i = sum(d[0] for d in data); tot_s = sum(d[1] for d in data)
    print('==', name[:80], 'instr', tot_i, 'samples', tot_s, 'sass lines', len(data))
    # cumulative by contiguous regions: print all lines with running index, marking top
    thr = sorted((d[0] for d in data), reverse=True)[min(top, len(data) - 1)]
    for k, d in enumerate(data):
        if d[0] >= thr and d[0] > 0:
            print('%5d %10d %5.1f%% samp %5.1f%% thr %5s  %s' % (k, d[0], 100. * d[0] / tot_i, 100. * d[1] / max(tot_s, 1), d[2][:5], d[3][:90]))
