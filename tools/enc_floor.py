#!/usr/bin/env python
"""Fixed-cost probe of ronk_match_encode: tiny problems, with/without L2 flush, graph replay."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ron_tensorflow_b200 import core, synth
from ron_tensorflow_b200.nets import ron_vgg_320
net = ron_vgg_320.RONNet(); aset = net.anchors((320, 320)).anchor_set
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
N = aset.N
def timeit(fn, do_flush, iters=12):
    ts = []
    for it in range(iters):
        if do_flush: flush.zero_()
        else: torch.cuda._sleep(200000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        if it >= 4: ts.append(a.elapsed_time(b) * 1e3)
    return sum(ts) / len(ts)
for B, glo, ghi in [(1, 1, 1), (1, 50, 50), (64, 1, 1), (64, 1, 50), (64, 50, 50), (256, 1, 50)]:
    boxes, labels, counts = synth.make_gt_batch(2, B, glo, ghi)
    d = [torch.from_numpy(x).cuda() for x in (boxes, labels, counts)]
    out = dict(labels=torch.empty((B, N), dtype=torch.int64, device='cuda'), loc=torch.empty((B, N, 4), device='cuda'),
               scores=torch.empty((B, N), device='cuda'))
    fn = lambda: core.match_encode(aset, d[0], d[1], d[2], 0.56, 0.3, out=out)
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            fn()
    torch.cuda.synchronize()
    print('B=%3d G=%2d..%2d  flush+call %.1f  noflush+call %.1f  flush+graph %.1f  noflush+graph %.1f us' % (
        B, glo, ghi, timeit(fn, True), timeit(fn, False), timeit(g.replay, True), timeit(g.replay, False)))
