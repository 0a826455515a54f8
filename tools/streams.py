#!/usr/bin/env python
"""Whole-job throughput of back-to-back steps on 1 / 2 / 3 / 4 CUDA streams (batches alternate between streams)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ron_tensorflow_b200 import core, synth
from ron_tensorflow_b200.nets import ron_vgg_320
net = ron_vgg_320.RONNet(); anchors = net.anchors((320, 320)); aset = anchors.anchor_set
N = aset.N

def run(step, nstreams, steps=48, warm=8):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    main = torch.cuda.current_stream()
    def loop(k0, k1):
        for k in range(k0, k1):
            with torch.cuda.stream(streams[k % nstreams]):
                step(k)
    for s in streams: s.wait_stream(main)
    loop(0, warm)
    for s in streams: main.wait_stream(s)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(main)
    for s in streams: s.wait_stream(main)
    loop(0, steps)
    for s in streams: main.wait_stream(s)
    b.record(main)
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / steps

B = 64
boxes, labels, counts = synth.make_gt_batch(2, B, 1, 50)
d = [torch.from_numpy(x).cuda() for x in (boxes, labels, counts)]
outs = [dict(labels=torch.empty((B, N), dtype=torch.int64, device='cuda'), loc=torch.empty((B, N, 4), device='cuda'),
             scores=torch.empty((B, N), device='cuda')) for _ in range(8)]          # 8 x 38 MB > L2
def enc(k): core.match_encode(aset, d[0], d[1], d[2], 0.56, 0.3, out=outs[k % 8])
for ns in ((1, 2, 3, 4) if 'post-only' not in sys.argv else ()):
    print('encode B=64  streams=%d  %.1f us/step' % (ns, run(enc, ns)))

PB = 256
ls = aset.layer_sizes
loc, pred, obj = synth.make_predictions(3000, PB, N, 21, hot=300)
dl = [torch.from_numpy(t).cuda() for t in synth.split_layers(loc, ls)]
dp = [torch.from_numpy(t).cuda() for t in synth.split_layers(pred, ls)]
do = [torch.from_numpy(t).cuda() for t in synth.split_layers(obj, ls)]
gb, gl, gc = synth.make_gt_batch(3, PB, 1, 12, g_max=12)
dg = [torch.from_numpy(x).cuda() for x in (gl, gb, gl * 0)]
def post(k):
    ns, nb = net.detect(dp, dl, do, 0.03, 0.01, 0.45, [0., 0., 1., 1.], 400, 200)
    core.tpfp_match(ns, nb, dg[0], dg[1], dg[2], 0.5)
for ns in ((1, 2, 3, 4, 6) if 'post-only' not in sys.argv else (1, 3)):
    print('post B=256   streams=%d  %.1f us/step' % (ns, run(post, ns, steps=24, warm=4)))
