#!/usr/bin/env python
"""BASELINE config 5 shape (crowded scenes): 81 classes, dense scores, up to 200 GT boxes; K = 400 and 10 000."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ron_tensorflow_b200 import core, synth
from ron_tensorflow_b200.nets import ron_vgg_320
net = ron_vgg_320.RONNet(); anchors = net.anchors((320, 320)); aset = anchors.anchor_set
N, C, B = aset.N, 81, int(sys.argv[1]) if len(sys.argv) > 1 else 16
loc, pred, obj = synth.make_predictions(5005, B, N, C, hot=2000, dense=True)
obj = np.maximum(obj, np.float32(0.05))
ls = aset.layer_sizes
dl = [torch.from_numpy(t).cuda() for t in synth.split_layers(loc, ls)]
dp = [torch.from_numpy(t).cuda() for t in synth.split_layers(pred, ls)]
do = [torch.from_numpy(t).cuda() for t in synth.split_layers(obj, ls)]
def timeit(fn, it=6):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
for K, M in ((400, 200), (10000, 200)):
    ms = timeit(lambda: net.detect(dp, dl, do, 0.03, 0.004, 0.45, [0., 0., 1., 1.], K, M))
    s, bx, _ = core.decode_select_topk(aset, dl, dp, do, 0.03, 0.004, [0., 0., 1., 1.], 0.03, K)
    print('post-process B=%d C=%d K=%d M=%d: %.2f ms/batch  (%.0f img/s), candidates/class ~%d' % (
        B, C, K, M, ms, B / ms * 1e3, int((s[0] > 0).sum(1).float().mean())))
boxes, labels, counts = synth.make_gt_batch(5, B, 120, 200, num_classes=C)
d = [torch.from_numpy(x).cuda() for x in (boxes, labels, counts)]
ms = timeit(lambda: core.match_encode(aset, d[0], d[1], d[2], 0.5, 0.3))
print('match+encode B=%d, 120-200 GT/image: %.3f ms/batch (%.0f img/s)' % (B, ms, B / ms * 1e3))
