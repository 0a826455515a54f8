#!/usr/bin/env python
"""Group an ncu source-page SASS listing into regions of equal execution count.
   python tools/sass_regions.py rep kernel-substr [min_pct]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
for blk in out.split('"Kernel Name",')[1:]:
    lines = blk.split('\n')
    if kern not in lines[0]:
        continue
    rows = list(csv.reader(io.StringIO('\n'.join(lines[1:]))))
    h = rows[0]
    iS, iN, iI, iT = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed'), h.index('Avg. Threads Executed')
    data = []
    for r in rows[1:]:
        if len(r) < len(h):
            continue
        try:
            data.append((int(r[iI] or 0), int(r[iN] or 0), float(r[iT] or 0), r[iS]))
        except ValueError:
            pass
    tot_i = sum(d[0] for d in data); tot_s = max(1, sum(d[1] for d in data))
    print('==', lines[0][:70], 'instr', tot_i, 'samples', tot_s)
    groups = []
    for k, d in enumerate(data):
        if groups and abs(groups[-1]['e'] - d[0]) <= 0.03 * max(d[0], 1):
            g = groups[-1]; g['n'] += 1; g['last'] = k; g['s'] += d[1]; g['i'] += d[0]; g['t'] += d[2]
            g['ops'].append(d[3].split()[0] if d[3].split() else '')
        else:
            groups.append({'first': k, 'last': k, 'e': d[0], 'n': 1, 's': d[1], 'i': d[0], 't': d[2], 'src': d[3], 'ops': [d[3].split()[0] if d[3].split() else '']})
    for g in groups:
        if 100. * g['i'] / tot_i >= minpct or 100. * g['s'] / tot_s >= 2 * minpct:
            ops = {}
            for o in g['ops']:
                o = o.split('.')[0]
                ops[o] = ops.get(o, 0) + 1
            top = ' '.join('%s:%d' % kv for kv in sorted(ops.items(), key=lambda kv: -kv[1])[:7])
            print('L%4d-%4d exec %9d x%4d = %5.1f%% instr, %5.1f%% samples, thr %4.1f | %s' % (
                g['first'], g['last'], g['e'], g['n'], 100. * g['i'] / tot_i, 100. * g['s'] / tot_s, g['t'] / g['n'], top))
