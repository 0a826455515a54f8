#!/usr/bin/env python
"""Event-timed post-process (detect + TP/FP) at one batch size: median of --iters single-stream steps, inputs > L2.
    python tools/post_time.py [--batch 256] [--iters 30] [--topk 400]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from ron_tensorflow_b200 import core, synth  # noqa: E402
from ron_tensorflow_b200.nets import ron_vgg_320  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=256)
ap.add_argument('--iters', type=int, default=30)
ap.add_argument('--topk', type=int, default=400)
ap.add_argument('--classes', type=int, default=21)
args = ap.parse_args()
net = ron_vgg_320.RONNet()
aset = net.anchors((320, 320)).anchor_set
B = args.batch
ls = aset.layer_sizes
loc, pred, obj = synth.make_predictions(3000, B, aset.N, args.classes, hot=300)
dl = [torch.from_numpy(t).cuda() for t in synth.split_layers(loc, ls)]
dp = [torch.from_numpy(t).cuda() for t in synth.split_layers(pred, ls)]
do = [torch.from_numpy(t).cuda() for t in synth.split_layers(obj, ls)]
gb, gl, gc = synth.make_gt_batch(3, B, 1, 12, g_max=12)
gbd, gld = torch.from_numpy(gb).cuda(), torch.from_numpy(gl).cuda()


def step():
    ns, nb = net.detect(dp, dl, do, 0.03, 0.01, 0.45, [0., 0., 1., 1.], args.topk, 200)
    core.tpfp_match(ns, nb, gld, gbd, gld * 0, 0.5)
    return ns


for _ in range(5):
    ns = step()
torch.cuda.synchronize()
ts = []
for _ in range(args.iters):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    step()
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
ts = np.array(ts)
print('post B=%d topk=%d: median %.1f us  p10 %.1f  min %.1f   kept %d' % (B, args.topk, np.median(ts), np.percentile(ts, 10), ts.min(), int((ns > 0).sum())))
