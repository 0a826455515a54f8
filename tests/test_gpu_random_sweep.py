"""Seeded random sweep of shapes and parameters against the oracle: small / ragged / odd sizes that
the fixed-size tests never hit (tiles that end mid-word for the bulk copies, a single anchor, one
class, K larger than the list, empty images, thresholds None / 0, every optional stage on and off)."""
import numpy as np
import pytest

from oracle import ron_oracle as O
from ron_tensorflow_b200 import synth
from _util import need_cuda, eq

pytestmark = pytest.mark.gpu


def _flat_anchors(rng, n):
    """random flat (y, x, h, w) anchors, some sticking out of the image"""
    yx = rng.uniform(-0.05, 1.05, size=(n, 2))
    hw = np.exp(rng.uniform(np.log(0.04), np.log(0.9), size=(n, 2)))
    return np.concatenate([yx, hw], 1).astype(np.float32)


@pytest.mark.parametrize('seed', range(12))
def test_encode_random_shapes(seed):
    need_cuda()
    from ron_tensorflow_b200 import core
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    n = int(rng.choice([1, 2, 31, 64, 65, 257, 1000, 3001]))
    yxhw = _flat_anchors(rng, n)
    border = rng.integers(0, 40, size=n).astype(np.int32) if seed % 3 else None
    aset = core.AnchorSet.flat((320, 320), yxhw, border)
    B = int(rng.choice([1, 2, 5, 70]))
    gmax = int(rng.choice([1, 3, 40, 130]))
    boxes, labels, counts = synth.make_gt_batch(50 + seed, B, 1, gmax, g_max=gmax)
    pos = float(rng.choice([0.5, 0.56, 0.7]))
    ign = float(rng.choice([0.3, 0.4, pos]))
    ib, gf = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
    r = core.match_encode(aset, boxes, labels, counts, pos, ign, ignore_between=ib, gt_max_first=gf,
                          want_matched=True, want_objness=True)
    # oracle tables for flat anchors: one corner trip (ssd_common.py:105-108), per-anchor borders
    y, x, h, w = (yxhw[:, k] for k in range(4))
    cor = np.stack([y - h / np.float32(2), x - w / np.float32(2), y + h / np.float32(2), x + w / np.float32(2)], -1)
    if border is not None:
        b = border.astype(np.float64)
        inside = (cor[:, 0] >= (-b / 320.).astype(np.float32)) & (cor[:, 1] >= (-b / 320.).astype(np.float32)) & \
                 (cor[:, 2] < ((320 + b) / 320.).astype(np.float32)) & (cor[:, 3] < ((320 + b) / 320.).astype(np.float32))
    else:
        inside = np.ones(n, bool)
    for i in range(B):
        o = O.encode_image(labels[i, :counts[i]], boxes[i, :counts[i]], yxhw, cor.astype(np.float32), inside, pos, ign,
                           ignore_between=ib, gt_max_first=gf)
        eq(r['matched'][i], o['matched'].astype(np.int32), 'matched seed %d img %d' % (seed, i))
        eq(r['labels'][i], o['labels'], 'labels')
        eq(r['scores'][i], o['scores'], 'scores')
        eq(r['loc'][i], o['loc'], 'loc')
        eq(r['objness'][i], o['objness'], 'objness')


@pytest.mark.parametrize('seed', range(16))
def test_postprocess_random_shapes(seed):
    need_cuda()
    import torch
    from ron_tensorflow_b200 import core
    rng = np.random.Generator(np.random.PCG64(2000 + seed))
    n = int(rng.choice([1, 3, 127, 128, 129, 255, 513, 2000, 4999]))
    C = int(rng.choice([2, 3, 5, 21, 22, 33]))
    B = int(rng.choice([1, 2, 3]))
    K = int(rng.choice([1, 7, 50, 400]))
    M = int(rng.choice([1, 5, 200]))
    yxhw = _flat_anchors(rng, n)
    aset = core.AnchorSet.flat((320, 320), yxhw, None)
    loc, pred, obj = synth.make_predictions(9000 + seed, B, n, C, hot=min(n, 40), dense=bool(seed % 2))
    use_obj = bool(rng.integers(0, 2))
    sel = [None, 0.0, 0.01, 0.2][int(rng.integers(0, 4))]
    clip = [0., 0., 1., 1.] if rng.integers(0, 2) else None
    minsize = 0.03 if rng.integers(0, 2) else None
    thr = float(rng.choice([0.3, 0.45, 0.6]))
    mode = 'min' if rng.integers(0, 2) else 'union'
    s, bx, ix = core.decode_select_topk(aset, [loc], [pred], [obj] if use_obj else None, 0.03, sel, clip, minsize, K,
                                        want_idx=True)
    ns, nb, ni = core.nms_batch(s.view(B * (C - 1), K), bx.view(B * (C - 1), K, 4), thr, M, mode, assume_sorted=True,
                                want_idx=True)
    for b in range(B):
        boxes = O.decode(loc[b], yxhw)
        p = pred[b]
        if use_obj:
            p = (obj[b] > np.float32(0.03)).astype(np.float32)[:, None] * p
        os_, ob, oi = O.select_topk_image(p, boxes, sel, K, clip, minsize)
        eq(ix[b], oi, 'top-k idx seed %d' % seed)
        eq(s[b], os_, 'top-k scores')
        eq(bx[b], ob, 'top-k boxes')
        on, onb, oni = O.nms_batch(os_, ob, thr, M, mode)
        eq(ns.view(B, C - 1, -1)[b], on[:, :M], 'nms scores')
        eq(nb.view(B, C - 1, -1, 4)[b], onb[:, :M], 'nms boxes')
        eq(ni.view(B, C - 1, -1)[b], oni[:, :M], 'nms positions')
