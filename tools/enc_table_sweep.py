#!/usr/bin/env python
"""Generic encode kernel: work-item table (0 / 1 / 2) against batch size and ground-truth count (CUDA events, L2 flushed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ron_tensorflow_b200 import core, synth
from ron_tensorflow_b200.nets import ron_vgg_320
net = ron_vgg_320.RONNet(); aset = net.anchors((320, 320)).anchor_set
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
os.environ['RONK_ENC_KERNEL'] = 'generic'
for (gmin, gmax) in ((1, 50), (120, 200), (40, 100)):
    for B in (8, 16, 32, 64, 96):
        boxes, labels, counts = synth.make_gt_batch(5, B, gmin, gmax, num_classes=81)
        d = [torch.from_numpy(x).cuda() for x in (boxes, labels, counts)]
        res = []
        for t in ('0', '1', '2'):
            os.environ['RONK_ENC_TABLE'] = t
            ts = []
            for it in range(14):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); core.match_encode(aset, d[0], d[1], d[2], 0.5, 0.3); b.record()
                torch.cuda.synchronize()
                if it >= 4: ts.append(a.elapsed_time(b) * 1e3)
            res.append(min(ts))
        print('GT %3d-%3d  B=%3d   table0 %.1f  table1 %.1f  table2 %.1f us   best %d' % (gmin, gmax, B, res[0], res[1], res[2], res.index(min(res))))
