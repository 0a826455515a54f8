#!/usr/bin/env python
"""Times the fused RON loss-mask kernel (CUDA events, L2 flushed) at batch 64."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ron_tensorflow_b200.nets import ron_vgg_320 as rv
rng = np.random.Generator(np.random.PCG64(1))
n = 64 * 21250
dev = 'cuda'
gc = torch.from_numpy(rng.choice(np.array([-1, 0, 0, 0, 0, 0, 0, 0, 3, 17], np.int64), size=n)).to(dev)
ob = torch.from_numpy(rng.random(n, dtype=np.float32)).to(dev)
r1 = torch.from_numpy(rng.random(n, dtype=np.float32)).to(dev)
r2 = torch.from_numpy(rng.random(n, dtype=np.float32)).to(dev)
lc = torch.from_numpy(rng.normal(0, 1, (n, 4)).astype(np.float32)).to(dev)
gl = torch.from_numpy(rng.normal(0, 1, (n, 4)).astype(np.float32)).to(dev)
flush = torch.empty(1 << 30, dtype=torch.uint8, device=dev)   # long enough for the host to get ahead
for name, fn in (('masks only', lambda: rv.ron_loss_masks(gc, ob, r1, r2)), ('masks + loc', lambda: rv.ron_loss_masks(gc, ob, r1, r2, localisations=lc, glocalisations=gl))):
    ts = []
    for it in range(14):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if it >= 4: ts.append(a.elapsed_time(b) * 1e3)
    print(name, '%.1f us' % (sum(ts) / len(ts)))
