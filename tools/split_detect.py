#!/usr/bin/env python
"""Does splitting one post-process call into batch halves on two streams shorten the call?  (scatter is HBM bound,
top-k / NMS latency bound: the halves' kernels can overlap.)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ron_tensorflow_b200 import core, synth
from ron_tensorflow_b200.nets import ron_vgg_320
net = ron_vgg_320.RONNet(); aset = net.anchors((320, 320)).anchor_set
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ls = aset.layer_sizes
loc, pred, obj = synth.make_predictions(3000, B, aset.N, 21, hot=300)
dl = [torch.from_numpy(t).cuda() for t in synth.split_layers(loc, ls)]
dp = [torch.from_numpy(t).cuda() for t in synth.split_layers(pred, ls)]
do = [torch.from_numpy(t).cuda() for t in synth.split_layers(obj, ls)]
gb, gl, gc = synth.make_gt_batch(3, B, 1, 12, g_max=12)
gbd, gld = torch.from_numpy(gb).cuda(), torch.from_numpy(gl).cuda()
gdd = gld * 0

def whole():
    ns, nb = net.detect(dp, dl, do, 0.03, 0.01, 0.45, [0., 0., 1., 1.], 400, 200)
    return core.tpfp_match(ns, nb, gld, gbd, gdd, 0.5)

def make_split(parts):
    streams = [torch.cuda.Stream() for _ in range(parts - 1)]
    cuts = [(B * i // parts, B * (i + 1) // parts) for i in range(parts)]
    def run():
        main = torch.cuda.current_stream()
        outs = []
        for i, (a, b) in enumerate(cuts):
            st = main if i == 0 else streams[i - 1]
            if st is not main: st.wait_stream(main)
            with torch.cuda.stream(st):
                ns, nb = net.detect([t[a:b] for t in dp], [t[a:b] for t in dl], [t[a:b] for t in do], 0.03, 0.01, 0.45, [0., 0., 1., 1.], 400, 200)
                outs.append(core.tpfp_match(ns, nb, gld[a:b], gbd[a:b], gdd[a:b], 0.5))
        for st in streams: main.wait_stream(st)
        return outs
    return run

def timeit(fn, graph, it=40):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    if graph:
        g = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream(); cap.wait_stream(torch.cuda.current_stream())
        with torch.cuda.graph(g, stream=cap): fn()
        torch.cuda.current_stream().wait_stream(cap)
        fn = g.replay
        fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(it):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))

for graph in (False, True):
    print('B=%d %s: whole %.1f us' % (B, 'graph' if graph else 'eager', timeit(whole, graph)), end='')
    for parts in (2, 4):
        print('   %d parts %.1f us' % (parts, timeit(make_split(parts), graph)), end='')
    print()
