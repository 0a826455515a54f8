#!/usr/bin/env python
"""Times ronk_match_encode over batch sizes / RONK_ENC_TABLE settings (CUDA events, L2 flushed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ron_tensorflow_b200 import core, synth
from ron_tensorflow_b200.nets import ron_vgg_320
net = ron_vgg_320.RONNet(); aset = net.anchors((320, 320)).anchor_set
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for B in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else '1,16,64,128,256').split(',')]:
    boxes, labels, counts = synth.make_gt_batch(2, B, 1, 50)
    d = [torch.from_numpy(x).cuda() for x in (boxes, labels, counts)]
    for split in (os.environ.get('VARIANTS', '0,1,2,auto').split(',')):
        if split == 'auto': os.environ.pop('RONK_ENC_TABLE', None)
        else: os.environ['RONK_ENC_TABLE'] = split
        ts = []
        for it in range(12):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); core.match_encode(aset, d[0], d[1], d[2], 0.56, 0.3); b.record()
            torch.cuda.synchronize()
            if it >= 4: ts.append(a.elapsed_time(b) * 1e3)
        print('B=%4d table=%4s  %.1f us (min %.1f)' % (B, split, sum(ts) / len(ts), min(ts)))
