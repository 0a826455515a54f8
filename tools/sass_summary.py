#!/usr/bin/env python
"""Per-kernel SASS evidence from the built library (no GPU needed): registers, instruction count and the counts of the
mnemonics that prove which hardware paths a kernel uses -- UBLKCP / SYNCS (1-D TMA bulk copies + mbarrier),
CREDUX / REDUX (warp reductions), MUFU.RCP (inline division), ATOMS / ATOMG / RED (atomics), LDS / STS, LDG / STG, BAR.
    python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'ron_tensorflow_b200', 'libronk.so')
WANT = ['UBLKCP', 'SYNCS', 'CREDUX', 'REDUX', 'MUFU.RCP', 'MUFU.EX2', 'ATOMS', 'ATOMG', 'RED.', 'LDS', 'STS', 'LDG', 'STG', 'LDL', 'STL',
        'BAR.', 'VOTE', 'SHFL', 'MATCH', 'DFMA', 'HMMA', 'UTC']
sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
res = subprocess.run(['cuobjdump', '-res-usage', LIB], capture_output=True, text=True).stdout
regs = {}
for m in re.finditer(r'Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)', res):
    regs[m.group(1)] = (int(m.group(2)), int(m.group(3)), int(m.group(4)))
cur, counts, total = None, collections.defaultdict(collections.Counter), collections.Counter()
for line in sass.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if cur and m:
        op = m.group(1)
        total[cur] += 1
        for w in WANT:
            if op.startswith(w):
                counts[cur][w] += 1
print('SASS summary of ron_tensorflow_b200/libronk.so (cuobjdump -sass / -res-usage, sm_100a); counts are static instructions')
print('%-46s %5s %6s %7s  %s' % ('kernel', 'regs', 'stack', 'instr', 'mnemonics'))
for k in sorted(total, key=lambda k: -total[k]):
    try:
        name = subprocess.run(['c++filt', k], capture_output=True, text=True).stdout.strip().split('(')[0]
    except Exception:
        name = k
    r = regs.get(k, (0, 0, 0))
    print('%-46s %5d %6d %7d  %s' % (name[-46:], r[0], r[1], total[k], '  '.join('%s=%d' % (w, counts[k][w]) for w in WANT if counts[k][w])))
