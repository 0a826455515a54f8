#!/usr/bin/env python
"""Tiny driver for ncu captures: runs one stage of the hot path a few times.
    ncu ... python tools/prof.py --stage encode --batch 256 --iters 3
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from ron_tensorflow_b200 import core, synth  # noqa: E402
from ron_tensorflow_b200.nets import ron_vgg_320  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--stage', default='encode')
ap.add_argument('--batch', type=int, default=256)
ap.add_argument('--iters', type=int, default=3)
ap.add_argument('--classes', type=int, default=21)
args = ap.parse_args()
net = ron_vgg_320.RONNet()
anchors = net.anchors((320, 320))
aset = anchors.anchor_set
B = args.batch
if args.stage == 'encode':
    boxes, labels, counts = synth.make_gt_batch(2, B, 1, 50)
    d = [torch.from_numpy(x).cuda() for x in (boxes, labels, counts)]
    for _ in range(args.iters):
        r = core.match_encode(aset, d[0], d[1], d[2], 0.56, 0.3)
    torch.cuda.synchronize()
    print('pos', int((r['labels'] > 0).sum()))
else:
    ls = aset.layer_sizes
    loc, pred, obj = synth.make_predictions(3000, B, aset.N, args.classes, hot=300)
    dl = [torch.from_numpy(t).cuda() for t in synth.split_layers(loc, ls)]
    dp = [torch.from_numpy(t).cuda() for t in synth.split_layers(pred, ls)]
    do = [torch.from_numpy(t).cuda() for t in synth.split_layers(obj, ls)]
    gb, gl, gc = synth.make_gt_batch(3, B, 1, 12, g_max=12)
    for _ in range(args.iters):
        ns, nb = net.detect(dp, dl, do, 0.03, 0.01, 0.45, [0., 0., 1., 1.], 400, 200)
        core.tpfp_match(ns, nb, gl, gb, gl * 0, 0.5)
    torch.cuda.synchronize()
    print('kept', int((ns > 0).sum()))
