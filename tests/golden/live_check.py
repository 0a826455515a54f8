#!/usr/bin/env python
"""Live cross-check of the oracle against the reference's OWN code on fresh random inputs (build container only:
needs /root/reference).  The fixtures under tests/golden/ pin fixed cases; this script draws new ones every run
(seed on the command line) and compares oracle/ron_oracle.py with the reference functions executed over the TF-1
shim -- and with the unmodified NumPy reference (nets/np_methods.py) where one exists.  Exit status 0 = all equal.

    python tests/golden/live_check.py [seed] [cases]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('RON_REFERENCE', '/root/reference')
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'tf1_shim'))
sys.path.insert(1, REF)
sys.path.insert(2, ROOT)

import tensorflow as tf  # noqa: E402  (the shim)
from nets import ron_vgg_320, ssd_common  # noqa: E402  (the reference)
import tf_extended as tfe  # noqa: E402  (the reference)
from oracle import ron_oracle as O  # noqa: E402

if not hasattr(np, 'bool'):
    np.bool = bool
from nets import np_methods  # noqa: E402  (the reference, pure NumPy)

T = tf.convert_to_tensor
npy = lambda x: x.a if isinstance(x, tf.Tensor) else np.asarray(x)


def same(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    ok = a.shape == b.shape and bool(np.all((a == b) | ((a != a) & (b != b))))
    if not ok:
        raise SystemExit('MISMATCH in %s' % what)


def boxes(rng, n, lo=0.02, hi=0.7):
    c = rng.uniform(0.05, 0.95, (n, 2))
    s = np.exp(rng.uniform(np.log(lo), np.log(hi), (n, 2)))
    b = np.concatenate([np.clip(c - s / 2, 0, 1), np.clip(c + s / 2, 0, 1)], 1).astype(np.float32)
    return b


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else int.from_bytes(os.urandom(4), 'little')
    cases = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    rng = np.random.Generator(np.random.PCG64(seed))
    net = ron_vgg_320.RONNet()
    anchors = net.anchors(net.params.img_shape)
    enc, cor, inside = O.encode_anchor_tables(O.anchors_all_layers(O.RON320), O.RON320.img_shape, O.RON320.allowed_borders)
    for k in range(cases):
        # ---- joint match + encode (ssd_common.py:337-414, 77-147, 27-75)
        G = int(rng.integers(1, 30))
        gb = boxes(rng, G)
        if k % 2:
            gb[rng.integers(0, G)] = gb[0]                                    # duplicated GT: ties
        gl = rng.integers(1, 21, G).astype(np.int64)
        pos, ign = float(rng.choice([0.5, 0.56, 0.7])), 0.3
        r = net.bboxes_encode(T(gl), T(gb), anchors, positive_threshold=pos, ignore_threshold=ign)
        mine = O.encode_image(gl, gb, enc, cor, inside, pos, ign)
        same(mine['labels'], np.concatenate([npy(t).reshape(-1) for t in r[0]]), 'encode labels')
        same(mine['loc'], np.concatenate([npy(t).reshape(-1, 4) for t in r[1]]), 'encode loc')
        same(mine['scores'], np.concatenate([npy(t).reshape(-1) for t in r[2]]), 'encode scores')
        # ---- NMS, both modes (tf_extended/bboxes.py:173-234)
        K = int(rng.integers(5, 120))
        b = boxes(rng, K, 0.05, 0.5)
        s = rng.uniform(0, 1, K).astype(np.float32)
        s[rng.integers(0, K, 3)] = s[0]
        for mode in ('min', 'union'):
            thr, M = float(rng.uniform(0.2, 0.7)), int(rng.integers(3, 80))
            rs, rb = tfe.bboxes_nms(T(s), T(b), nms_threshold=thr, keep_top_k=M, mode=mode)
            ms, mb, _ = O.nms(s, b, thr, M, mode)
            same(ms, npy(rs), 'nms scores ' + mode); same(mb, npy(rb), 'nms boxes ' + mode)
        # ---- TP/FP matching of one class (tf_extended/bboxes.py:316-404)
        Gm = int(rng.integers(1, 8))
        gtb, gtl = boxes(rng, Gm, 0.1, 0.5), rng.integers(1, 4, Gm).astype(np.int64)
        gtd = (rng.uniform(size=Gm) < 0.3).astype(np.int64)
        D = int(rng.integers(1, 30))
        db = (gtb[rng.integers(0, Gm, D)] + rng.normal(0, 0.04, (D, 4))).astype(np.float32)
        ds = np.sort(rng.uniform(0, 1, D).astype(np.float32))[::-1].copy()
        n, tp, fp = tfe.bboxes_matching(2, T(ds), T(db), T(gtl), T(gtb), T(gtd), matching_threshold=0.5)
        mn, mtp, mfp = O.bboxes_matching(2, ds, db, gtl, gtb, gtd, 0.5)
        same(mn, npy(n), 'matching n_gt'); same(mtp, npy(tp), 'matching tp'); same(mfp, npy(fp), 'matching fp')
        # ---- the NumPy twin, unmodified reference (nets/np_methods.py)
        n = int(rng.integers(20, 400))
        pb, pc = boxes(rng, n, 0.05, 0.4), rng.integers(1, 5, n).astype(np.int64)
        ps = np.sort(rng.permutation(n).astype(np.float32) / n)[::-1].copy()
        thr = float(rng.uniform(0.2, 0.6))
        same(O.np_nms(pc, ps, pb, thr)[2], np_methods.bboxes_nms(pc, ps, pb, nms_threshold=thr)[2], 'np_methods nms')
        pred = rng.dirichlet(np.ones(6), n).astype(np.float32)
        for sel in (None, 0.3):
            a = O.np_select_layer(pred, pb, sel)
            bb = np_methods.ssd_bboxes_select_layer(pred[None], pb[None], None, select_threshold=sel, decode=False)
            for x, y, w in zip(a, bb, ('classes', 'scores', 'boxes')):
                same(x, y, 'np_methods select ' + w)
    # ---- the eval post-process chain of eval_ron_network.py:226-236 on one image (RONNet.detected_bboxes), once per run
    from ron_tensorflow_b200 import synth
    p = net.params
    layer_sizes = [a[0].shape[0] * a[0].shape[1] * a[2].shape[0] for a in anchors]
    apc = [a[2].shape[0] for a in anchors]
    hot = int(rng.integers(20, 200))
    loc, pred, obj = synth.make_predictions(int(rng.integers(1, 1 << 30)), 1, sum(layer_sizes), p.num_classes, hot=hot)
    split = lambda x: [T(t) for t in synth.split_layers(x, layer_sizes, p.feat_shapes, apc)]
    dec = net.bboxes_decode(split(loc), anchors)
    filt = [tf.cast(tf.greater(o, 0.03), tf.float32) * pr for o, pr in zip(split(obj[..., None]), split(pred))]
    K, M, thr = int(rng.integers(20, 80)), int(rng.integers(5, 40)), float(rng.uniform(0.3, 0.6))
    rs, rb = net.detected_bboxes(filt, dec, select_threshold=0.01, nms_threshold=thr, clipping_bbox=[0., 0., 1., 1.],
                                 top_k=K, keep_top_k=M)
    mine = O.detected_bboxes_image(pred[0], loc[0], O.flat_decode_anchors(O.anchors_all_layers(O.RON320)), obj[0], 0.03,
                                   0.01, thr, [0., 0., 1., 1.], K, M)
    classes = sorted(rs.keys())
    same(mine['scores'], np.stack([npy(rs[c])[0] for c in classes]), 'detected_bboxes scores')
    same(mine['boxes'], np.stack([npy(rb[c])[0] for c in classes]), 'detected_bboxes boxes')
    print('live check ok: seed %d, %d cases' % (seed, cases))


if __name__ == '__main__':
    main()
