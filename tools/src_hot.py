#!/usr/bin/env python
"""Per CUDA source line instruction / stall-sample totals from an ncu report.
   python tools/src_hot.py rep kernel-substr [top]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, cur_fn, hdr = None, None, None
agg = {}
seen_fn = set()
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name':
        cur_fn = r[1]; continue
    if r[0] == 'Line No':
        hdr = r; iI = hdr.index('Instructions Executed'); iN = hdr.index('# Samples'); continue
    if hdr is None or cur_fn is None or kern not in cur_fn:
        continue
    if r[0] != '' and r[0].isdigit():
        try:
            key = (cur_file, int(r[0]), r[1].strip()[:100])
            a = agg.setdefault(key, [0, 0])
            a[0] += int(r[iI] or 0); a[1] += int(r[iN] or 0)
        except (ValueError, IndexError):
            pass
ti = sum(a[0] for a in agg.values()); ts = max(1, sum(a[1] for a in agg.values()))
print('total instr', ti, 'samples', ts, '(all captured launches of the kernel summed)')
for (f, ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%5.1f%% instr %5.1f%% samp  %s:%d  %s' % (100. * a[0] / ti, 100. * a[1] / ts, f, ln, src))
