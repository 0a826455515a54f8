#!/usr/bin/env python
"""Per-CTA timeline of match_encode_kernel (start / end globaltimer + SM id) from a profiling build of the
library (-DRONK_ENC_TRACE, built into tools/_trace/): ramp, per-SM busy time, tail.
   python tools/enc_trace.py --build          (where nvcc is; the .so travels with the snapshot)
   python tools/enc_trace.py [batch]          (on the GPU box)"""
import os, sys, subprocess, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, 'tools', '_trace', 'libronk.so')
from ron_tensorflow_b200 import build as B
if '--build' in sys.argv:
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ['nvcc'] + B.NVCC_FLAGS + ['-DRONK_ENC_TRACE', '-I', os.path.join(ROOT, 'include'), '-o', OUT] + \
          [os.path.join(B.CSRC, s) for s in B.SOURCES]
    print(' '.join(cmd)); subprocess.check_call(cmd); sys.exit(0)
import numpy as np, torch
from ron_tensorflow_b200 import _ffi
_ffi.LIB_PATH = OUT
from ron_tensorflow_b200 import core, synth
from ron_tensorflow_b200.nets import ron_vgg_320
L = _ffi.lib()
setter = getattr(ctypes.CDLL(OUT), 'ronk_debug_set_enc_trace' if os.environ.get('RONK_ENC_KERNEL', 'grid').startswith('ge') else 'ronk_debug_set_enc_trace_grid')
setter.argtypes = [ctypes.c_void_p]
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
aset = ron_vgg_320.RONNet().anchors((320, 320)).anchor_set
boxes, labels, counts = synth.make_gt_batch(2, batch, 1, 50)
d = [torch.from_numpy(x).cuda() for x in (boxes, labels, counts)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for _ in range(3): core.match_encode(aset, d[0], d[1], d[2], 0.56, 0.3)
nmax = 65536 * 4
tr = torch.zeros(nmax * 8, dtype=torch.int64, device='cuda')
flush.zero_(); torch.cuda.synchronize()
setter(tr.data_ptr())
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); core.match_encode(aset, d[0], d[1], d[2], 0.56, 0.3); b.record(); torch.cuda.synchronize()
setter(None)
t = tr.cpu().numpy().reshape(-1, 8)
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
st, sm, en = (t[:, 0] - t0) / 1e3, t[:, 1], (t[:, 2] - t0) / 1e3
print('batch %d: %d CTAs, event time %.1f us, first start 0, last start %.1f us, last end %.1f us' % (batch, len(t), a.elapsed_time(b) * 1e3, st.max(), en.max()))
dur = en - st
print('CTA duration us: mean %.2f  p50 %.2f  p90 %.2f  p99 %.2f  max %.2f' % (dur.mean(), np.median(dur), np.percentile(dur, 90), np.percentile(dur, 99), dur.max()))
order = np.argsort(st)
print('start time of the k-th CTA: ' + '  '.join('%d:%.1f' % (k, st[order[k]]) for k in (0, 147, 591, 1183, 1500, 2000, 2500, 3000, len(t) - 1) if k < len(t)))
sms = np.unique(sm)
last_end = np.array([en[sm == s].max() for s in sms]); first_start = np.array([st[sm == s].min() for s in sms])
print('%d SMs: first start min/mean/max %.1f/%.1f/%.1f us; last end min/mean/max %.1f/%.1f/%.1f us' % (
    len(sms), first_start.min(), first_start.mean(), first_start.max(), last_end.min(), last_end.mean(), last_end.max()))
# resident CTAs over time
grid = np.linspace(0, en.max(), 41)
res = [(int(((st <= x) & (en > x)).sum())) for x in grid]
print('resident CTAs at t (us): ' + ' '.join('%.0f:%d' % (x, r) for x, r in zip(grid, res)))
late = order[-64:]
print('last 64 CTAs started: duration mean %.2f max %.2f us' % (dur[late].mean(), dur[late].max()))
heavy = np.argsort(-dur)[:10]
print('10 longest CTAs: ' + ' '.join('(start %.1f dur %.1f)' % (st[i], dur[i]) for i in heavy))
ph = (t[:, [3, 4, 5, 6, 2]] - t[:, [0, 3, 4, 5, 6]]) / 1e3
names = ['load+stage', 'sweep', 'output', 'fence+counter', 'force/exit']
if not os.environ.get('RONK_ENC_KERNEL', 'grid').startswith('ge'):
    print('sweep of warp 0 vs whole CTA (us): %.2f / %.2f' % (((t[:, 7] - t[:, 3]) / 1e3).mean(), ((t[:, 4] - t[:, 3]) / 1e3).mean()))
for sel, what in ((slice(None), 'all CTAs'), (late, 'last 64 started'), (order[:1184], 'first wave'), (order[1500:2500], 'middle')):
    print('%-16s ' % what + '  '.join('%s %.2f' % (n, v) for n, v in zip(names, ph[sel].mean(0))) + '   (mean us per phase)')
