import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from ron_tensorflow_b200 import core, synth
from ron_tensorflow_b200.nets import ron_vgg_320
net = ron_vgg_320.RONNet(); aset = net.anchors((320, 320)).anchor_set
N, C, B = aset.N, 81, 16
loc, pred, obj = synth.make_predictions(5005, B, N, C, hot=2000, dense=True)
obj = np.maximum(obj, np.float32(0.05))
ls = aset.layer_sizes
dl = [torch.from_numpy(t).cuda() for t in synth.split_layers(loc, ls)]
dp = [torch.from_numpy(t).cuda() for t in synth.split_layers(pred, ls)]
do = [torch.from_numpy(t).cuda() for t in synth.split_layers(obj, ls)]
K = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
for _ in range(3):
    ns, nb = net.detect(dp, dl, do, 0.03, 0.004, 0.45, [0., 0., 1., 1.], K, 200)
torch.cuda.synchronize()
kept = (ns > 0).sum(-1).float()
print('kept per segment: mean %.1f min %d max %d; segments full (200): %d of %d' % (kept.mean().item(), kept.min().item(), kept.max().item(), int((kept == 200).sum()), kept.numel()))
s, bx, idx = core.decode_select_topk(aset, dl, dp, do, 0.03, 0.004, [0., 0., 1., 1.], 0.03, K)
# how deep does NMS walk: position of the last kept candidate
pos = []
sc = s.reshape(-1, K); kn = ns.reshape(-1, 200)
for seg in range(0, sc.shape[0], 97):
    last = kn[seg][kn[seg] > 0][-1]
    pos.append(int((sc[seg] >= last).sum()))
print('position of the last kept candidate (sampled segments):', pos[:14], 'mean', np.mean(pos))
