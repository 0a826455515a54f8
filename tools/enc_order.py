#!/usr/bin/env python
"""Does the dispatch order of the images matter?  Same batch, original order vs sorted by GT count."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ron_tensorflow_b200 import core, synth
from ron_tensorflow_b200.nets import ron_vgg_320
net = ron_vgg_320.RONNet(); aset = net.anchors((320, 320)).anchor_set
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
N = aset.N
for B in (32, 64, 128, 256):
    boxes, labels, counts = synth.make_gt_batch(2, B, 1, 50)
    for name, order in (('orig', np.arange(B)), ('G desc', np.argsort(-counts, kind='stable')), ('G asc', np.argsort(counts, kind='stable'))):
        d = [torch.from_numpy(np.ascontiguousarray(x[order])).cuda() for x in (boxes, labels, counts)]
        out = dict(labels=torch.empty((B, N), dtype=torch.int64, device='cuda'), loc=torch.empty((B, N, 4), device='cuda'), scores=torch.empty((B, N), device='cuda'))
        ts = []
        for it in range(14):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); core.match_encode(aset, d[0], d[1], d[2], 0.56, 0.3, out=out); b.record()
            torch.cuda.synchronize()
            if it >= 4: ts.append(a.elapsed_time(b) * 1e3)
        print('B=%3d %-7s %.1f us' % (B, name, sum(ts) / len(ts)))
