"""nets/np_methods.py, the NumPy post-process of the reference's notebooks.
The fixture comes from the UNMODIFIED reference functions executed here (pure NumPy;
tests/golden/make_golden.py: gen_np_methods): select -> clip -> sort -> class-aware NMS -> resize,
plus jaccard / intersection."""
import hashlib

import numpy as np
import pytest

from oracle import ron_oracle as O
from ron_tensorflow_b200 import synth
from _util import need_cuda, eq, close

LS = [38 * 38 * 4, 19 * 19 * 6, 10 * 10 * 6, 5 * 5 * 6, 3 * 3 * 4, 1 * 1 * 4]
REF = np.array([0.05, 0.1, 0.9, 0.95], np.float32)


def _sha(*arrs):
    return np.frombuffer(hashlib.sha256(b''.join(np.ascontiguousarray(a).tobytes() for a in arrs)).digest(), np.uint8)


def _inputs(g, tag):
    seed, thr, nms_thr = [int(v) for v in g[tag + '_cfg']]
    loc, pred, _ = synth.make_predictions(seed, 1, sum(LS), 21, hot=120)
    if hashlib.sha256(pred.tobytes() + loc.tobytes()).digest() != g[tag + '_in_pred_sha'].tobytes():
        pytest.skip('numpy Generator stream differs from the one that made the fixture')
    return loc[0], pred[0], (None if thr < 0 else thr / 1000.), nms_thr / 1000.


def _decoded_by_oracle(loc):
    """The boxes the fixture's exact stages started from: the oracle's decode (correctly rounded exp); the
    reference's own np.exp decode differs in the last bit depending on the NumPy build."""
    return O.decode(loc, O.flat_decode_anchors(O.anchors_all_layers(O.SSD300)))


def _chain(sel, clip, sort, nms, resize, jac, inter, pred, boxes, thr, nms_thr):
    c, s, b = sel(pred, boxes, thr)
    r = {'sel': (c, s, b)}
    b = clip(REF, b)
    r['clipped'] = b
    c, s, b = sort(c, s, b, 400)
    r['sort'] = (c, s, b)
    c, s, b = nms(c, s, b, nms_thr)
    r['nms'] = (c, s, b)
    r['resized'] = resize(REF, b)
    r['jaccard'] = jac(b[0], b)
    r['intersection'] = inter(b[0], b)
    return r


def _check(r, g, tag, to_np):
    c, s, b = (to_np(x) for x in r['sel'])
    assert c.shape[0] == int(g[tag + '_sel_count'][0])
    assert np.array_equal(_sha(c.astype(np.int64), s.astype(np.float32), b.astype(np.float32)), g[tag + '_sel_sha']), 'selection'
    assert np.array_equal(_sha(to_np(r['clipped']).astype(np.float32)), g[tag + '_clipped_sha']), 'clip'
    for stage in ('sort', 'nms'):
        c, s, b = (to_np(x) for x in r[stage])
        eq(c.astype(np.int16), g['%s_%s_classes' % (tag, stage)], stage + ' classes')
        eq(s, g['%s_%s_scores' % (tag, stage)], stage + ' scores')
        eq(b, g['%s_%s_boxes' % (tag, stage)], stage + ' boxes')
    eq(to_np(r['resized']), g[tag + '_resized'], 'resized')
    eq(to_np(r['jaccard']), g[tag + '_jaccard'], 'jaccard')
    eq(to_np(r['intersection']), g[tag + '_intersection'], 'intersection')


@pytest.mark.parametrize('tag', ['a', 'b', 'c'])
def test_oracle_matches_reference_np_methods(golden, tag):
    g = golden('np_methods')
    loc, pred, thr, nms_thr = _inputs(g, tag)
    boxes = _decoded_by_oracle(loc)
    if tag == 'a':
        close(boxes[::7], g['a_decoded_7th'], 1e-5, 'decode (correctly rounded exp vs the reference np.exp boxes)')
    sort = lambda c, s, b, k: tuple(x[O.topk_stable(s, min(k, s.shape[0]))] for x in (c, s, b))
    inter = lambda ref, b: (np.maximum(np.minimum(ref[2], b[:, 2]) - np.maximum(ref[0], b[:, 0]), np.float32(0)) *
                            np.maximum(np.minimum(ref[3], b[:, 3]) - np.maximum(ref[1], b[:, 1]), np.float32(0))) / \
                           ((ref[2] - ref[0]) * (ref[3] - ref[1]))
    r = _chain(O.np_select_layer, O.np_clip, sort, O.np_nms, O.bboxes_resize, O.np_jaccard, inter, pred, boxes, thr, nms_thr)
    _check(r, g, tag, np.asarray)


@pytest.mark.gpu
@pytest.mark.parametrize('tag', ['a', 'b', 'c'])
def test_cuda_matches_reference_np_methods(golden, tag):
    need_cuda()
    from ron_tensorflow_b200.nets import np_methods as M
    g = golden('np_methods')
    loc, pred, thr, nms_thr = _inputs(g, tag)
    boxes = _decoded_by_oracle(loc)
    sel = lambda p, b, t: M.ssd_bboxes_select_layer(p, b, None, select_threshold=t, decode=False)
    r = _chain(sel, M.bboxes_clip, M.bboxes_sort, M.bboxes_nms, M.bboxes_resize, M.bboxes_jaccard, M.bboxes_intersection,
               pred, boxes, thr, nms_thr)
    _check(r, g, tag, lambda x: x.detach().cpu().numpy())
    if tag == 'a':
        # decode through the drop-in (per layer, from the (y, x, h, w) anchor tuples): correctly rounded exp,
        # within 1e-5 of the reference's np.exp boxes
        anchors = O.anchors_all_layers(O.SSD300)
        off, got = 0, []
        for a_l, n in zip(anchors, LS):
            got.append(M.ssd_bboxes_decode(loc[off:off + n].reshape(-1, a_l[2].size, 4), a_l).reshape(-1, 4))
            off += n
        import torch
        dec = torch.cat(got).cpu().numpy()
        eq(dec, _decoded_by_oracle(loc), 'decode vs oracle')
        close(dec[::7], g['a_decoded_7th'], 1e-5, 'decode vs np.exp boxes')


@pytest.mark.gpu
def test_cuda_np_nms_degenerate_and_large():
    """NaN overlaps (identical empty boxes) drop same-class boxes, other classes survive; 3 000 boxes against the oracle."""
    need_cuda()
    from ron_tensorflow_b200.nets import np_methods as M
    b = np.array([[.1, .1, .1, .1], [.1, .1, .1, .1], [.1, .1, .1, .1], [.2, .2, .6, .6], [.2, .2, .6, .6]], np.float32)
    c = np.array([3, 3, 4, 5, 5], np.int64)
    s = np.array([.9, .8, .7, .6, .5], np.float32)
    rc, rs, rb = O.np_nms(c, s, b, 0.45)
    gc, gs, gb = M.bboxes_nms(c, s, b, 0.45)
    assert rc.tolist() == [3, 4, 5]
    eq(gc, rc, 'classes'); eq(gs, rs, 'scores'); eq(gb, rb, 'boxes')
    rng = np.random.Generator(np.random.PCG64(9))
    n = 3000
    ctr = rng.uniform(0.2, 0.8, (n, 2)); sz = rng.uniform(0.05, 0.3, (n, 2))
    b = np.concatenate([ctr - sz / 2, ctr + sz / 2], 1).astype(np.float32)
    c = rng.integers(1, 4, n).astype(np.int64)
    s = np.sort(rng.random(n).astype(np.float32))[::-1].copy()
    rc, rs, rb = O.np_nms(c, s, b, 0.3)
    gc, gs, gb = M.bboxes_nms(c, s, b, 0.3)
    assert 10 < rc.shape[0] < n
    eq(gc, rc, 'classes'); eq(gs, rs, 'scores'); eq(gb, rb, 'boxes')
    gc, gs, gb = M.bboxes_nms(c[:0], s[:0], b[:0])
    assert gc.shape == (0,) and gb.shape == (0, 4)
