"""ron_eval.py single-image post-process variant: flaten_predict ->
clip -> filter_boxes -> class-agnostic tf_bboxes_nms (or the per-class tf_bboxes_nms_by_class / _v1) ->
bboxes_resize.  The golden fixture was made
by the reference's own function source executed over the TF-1 shim (tests/golden/make_golden.py)."""
import hashlib

import numpy as np
import pytest

from oracle import ron_oracle as O
from ron_tensorflow_b200 import synth
from _util import need_cuda, eq

LS = [250, 1000, 4000, 16000]
FS = [(5, 5), (10, 10), (20, 20), (40, 40)]
SEL, OBJ, NMS = 0.02, 0.03, 0.4


def _inputs(g, tag):
    seed, hot, keep, union, ih, iw = [int(v) for v in g[tag + '_cfg']]
    loc, pred, obj = synth.make_predictions(seed, 1, 21250, 21, hot=hot)
    if hashlib.sha256(pred.tobytes()).digest() != g[tag + '_in_pred_sha'].tobytes():
        pytest.skip('numpy Generator stream differs from the one that made the fixture')
    return loc, pred, obj, keep, ('union' if union else 'min'), (ih, iw)


@pytest.mark.parametrize('tag', ['a', 'b'])
def test_oracle_matches_reference_functions(golden, tag):
    g = golden('ron_eval')
    loc, pred, obj, keep, mode, img = _inputs(g, tag)
    dec = O.flat_decode_anchors(O.anchors_all_layers(O.RON320))
    boxes = O.decode(loc[0], dec)
    s, l, b = O.flaten_predict([pred[0]], [obj[0]], [boxes], OBJ)
    assert np.array_equal(s, g[tag + '_flat_scores']) and np.array_equal(l, g[tag + '_flat_labels'])
    assert np.array_equal(b, g[tag + '_flat_boxes'])
    b = O.clip_boxes([0., 0., 1., 1.], b)
    s, l, b = O.filter_boxes(s, l, b, 0.03, img, [320., 320.])
    assert np.array_equal(s, g[tag + '_filt_scores']) and np.array_equal(b, g[tag + '_filt_boxes'])
    cs, cl, cb = O.bboxes_nms_by_class(s, l, b, SEL, NMS, 10, mode)
    assert np.array_equal(cs, g[tag + '_bycls_scores']) and np.array_equal(cl, g[tag + '_bycls_labels'])
    assert np.array_equal(cb, g[tag + '_bycls_boxes'])
    cs, cl, cb = O.bboxes_nms_by_class_v1(s, l, b, SEL, 21, NMS, 12, mode)
    assert np.array_equal(cs, g[tag + '_v1_scores']) and np.array_equal(cl, g[tag + '_v1_labels'])
    assert np.array_equal(cb, g[tag + '_v1_boxes'])
    s, l, b = O.bboxes_nms_agnostic(s, l, b, SEL, NMS, keep, mode)
    assert np.array_equal(s, g[tag + '_nms_scores']) and np.array_equal(l, g[tag + '_nms_labels'])
    assert np.array_equal(b, g[tag + '_nms_boxes'])
    assert np.array_equal(O.bboxes_resize([0.1, 0.05, 0.9, 0.95], b), g[tag + '_resized'])


@pytest.mark.gpu
@pytest.mark.parametrize('tag', ['a', 'b'])
def test_cuda_matches_reference_functions(golden, tag):
    need_cuda()
    from ron_tensorflow_b200 import ron_eval
    from ron_tensorflow_b200.nets import ron_vgg_320
    import ron_tensorflow_b200.tf_extended as tfe
    g = golden('ron_eval')
    loc, pred, obj, keep, mode, img = _inputs(g, tag)
    net = ron_vgg_320.RONNet()
    anchors = net.anchors(net.params.img_shape)
    P = synth.split_layers(pred, LS, FS, [10] * 4)
    Ob = synth.split_layers(obj[..., None], LS, FS, [10] * 4)
    Lc = synth.split_layers(loc, LS, FS, [10] * 4)
    ron_eval.FLAGS.select_threshold, ron_eval.FLAGS.objectness_thres = SEL, OBJ
    boxes = net.bboxes_decode(Lc, anchors)
    s, l, b = ron_eval.flaten_predict(P, Ob, boxes)
    eq(s, g[tag + '_flat_scores'], 'flat scores'); eq(l, g[tag + '_flat_labels'], 'flat labels'); eq(b, g[tag + '_flat_boxes'], 'flat boxes')
    b = tfe.bboxes_clip([0., 0., 1., 1.], b)
    s, l, b = ron_eval.filter_boxes(s, l, b, 0.03, img, [320., 320.])
    eq(s, g[tag + '_filt_scores'], 'filtered scores'); eq(l, g[tag + '_filt_labels'], 'filtered labels'); eq(b, g[tag + '_filt_boxes'], 'filtered boxes')
    cs, cl, cb = ron_eval.tf_bboxes_nms_by_class(s, l, b, nms_threshold=NMS, keep_top_k=10, mode=mode)
    eq(cs, g[tag + '_bycls_scores'], 'by-class scores'); eq(cl, g[tag + '_bycls_labels'], 'by-class labels'); eq(cb, g[tag + '_bycls_boxes'], 'by-class boxes')
    cs, cl, cb = ron_eval.tf_bboxes_nms_by_class_v1(s, l, b, nms_threshold=NMS, keep_top_k=12, mode=mode)
    eq(cs, g[tag + '_v1_scores'], 'v1 scores'); eq(cl, g[tag + '_v1_labels'], 'v1 labels'); eq(cb, g[tag + '_v1_boxes'], 'v1 boxes')
    s, l, b = ron_eval.tf_bboxes_nms(s, l, b, nms_threshold=NMS, keep_top_k=keep, mode=mode)
    eq(s, g[tag + '_nms_scores'], 'nms scores'); eq(l, g[tag + '_nms_labels'], 'nms labels'); eq(b, g[tag + '_nms_boxes'], 'nms boxes')
    eq(tfe.bboxes_resize([0.1, 0.05, 0.9, 0.95], b), g[tag + '_resized'], 'resized')


@pytest.mark.gpu
def test_cuda_vs_oracle_dense_and_empty():
    """Dense scores (thousands of survivors: compaction across many tiles, sort of ~8k boxes) and the
    empty case (nothing passes: every stage must return zero-length tensors)."""
    need_cuda()
    import torch
    from ron_tensorflow_b200 import ron_eval
    from ron_tensorflow_b200.nets import ron_vgg_320
    net = ron_vgg_320.RONNet()
    anchors = net.anchors(net.params.img_shape)
    dec = O.flat_decode_anchors(O.anchors_all_layers(O.RON320))
    loc, pred, obj = synth.make_predictions(404, 1, 21250, 21, hot=300, dense=True)
    boxes = O.decode(loc[0], dec)
    P = synth.split_layers(pred, LS); Ob = synth.split_layers(obj[..., None], LS); Bx = synth.split_layers(boxes[None], LS)
    ron_eval.FLAGS.select_threshold, ron_eval.FLAGS.objectness_thres = 0.002, 0.03
    s, l, b = ron_eval.flaten_predict(P, Ob, [torch.from_numpy(t) for t in Bx])
    os_, ol, ob = O.flaten_predict([pred[0]], [obj[0]], [boxes], 0.03)
    assert os_.shape[0] > 5000
    eq(s, os_, 'scores'); eq(l, ol, 'labels'); eq(b, ob, 'boxes')
    for mode, keep in (('min', 40), ('union', 7)):               # per-class variants on the ~8k dense boxes
        cs, cl, cb = ron_eval.tf_bboxes_nms_by_class(s, l, b, nms_threshold=0.45, keep_top_k=keep, mode=mode)
        rs, rl, rb = O.bboxes_nms_by_class(os_, ol, ob, 0.002, 0.45, keep, mode)
        eq(cs, rs, 'by-class scores'); eq(cl, rl, 'by-class labels'); eq(cb, rb, 'by-class boxes')
        cs, cl, cb = ron_eval.tf_bboxes_nms_by_class_v1(s, l, b, nms_threshold=0.45, keep_top_k=keep, mode=mode)
        rs, rl, rb = O.bboxes_nms_by_class_v1(os_, ol, ob, 0.002, 21, 0.45, keep, mode)
        eq(cs, rs, 'v1 scores'); eq(cl, rl, 'v1 labels'); eq(cb, rb, 'v1 boxes')
    s, l, b = ron_eval.tf_bboxes_nms(s, l, b, nms_threshold=0.45, keep_top_k=300, mode='union')
    os_, ol, ob = O.bboxes_nms_agnostic(os_, ol, ob, 0.002, 0.45, 300, 'union')
    eq(s, os_, 'nms scores'); eq(l, ol, 'nms labels'); eq(b, ob, 'nms boxes')
    ron_eval.FLAGS.objectness_thres = 2.0                       # nothing passes
    s, l, b = ron_eval.flaten_predict(P, Ob, [torch.from_numpy(t) for t in Bx])
    assert s.shape == (0, 21) and l.shape == (0,) and b.shape == (0, 4)
    s, l, b = ron_eval.filter_boxes(s, l, b, 0.03, (375, 500), [320., 320.])
    for fn in (ron_eval.tf_bboxes_nms_by_class_v1, ron_eval.tf_bboxes_nms):
        rs, rl, rb = fn(s, l, b)
        assert rs.shape == (0,) and rl.shape == (0,) and rb.shape == (0, 4)
    rs, rl, rb = ron_eval.tf_bboxes_nms_by_class(s, l, b)      # n < 1: the inputs come back unchanged (:291)
    assert rs.shape == (0, 21) and rl.shape == (0,) and rb.shape == (0, 4)


@pytest.mark.gpu
def test_more_boxes_than_the_shared_memory_sort_takes():
    """The reference sorts with tf.nn.top_k(k = number of boxes) (ron_eval.py:155 / :217 / :301), which has no bound;
    beyond 16 384 boxes the CUDA path switches to the global-memory radix sort (ronk_sort_rows).  (a) core.sort_topk
    on rows of 40 000 scores with heavy ties, negative values and +-0 against the oracle's stable top-k; (b) the three
    NMS variants on all ~21 000 RON-320 boxes of one image (thresholds low enough that nearly every box passes) against
    the oracle; (c) core.nms_batch on unsorted rows longer than 16 384."""
    need_cuda()
    import torch
    from ron_tensorflow_b200 import core, ron_eval
    rng = np.random.Generator(np.random.PCG64(77))
    S, N = 3, 40000
    sc = rng.normal(0., 1., size=(S, N)).astype(np.float32)
    sc[0] = np.round(sc[0], 1)                                       # ~60 distinct values: long tie runs
    sc[1, ::7] = 0.
    sc[1, 3::11] = -0.
    bx = rng.uniform(0, 1, size=(S, N, 4)).astype(np.float32)
    for K in (N, 20000):
        ss, sb, si = core.sort_topk(torch.from_numpy(sc), torch.from_numpy(bx), K, want_idx=True)
        for r in range(S):
            o = O.topk_stable(sc[r], K)
            eq(si[r], o.astype(np.int32), 'order of row %d' % r)
            eq(ss[r], sc[r][o], 'scores'); eq(sb[r], bx[r][o], 'boxes')
    # (c) unsorted NMS rows
    b2 = np.concatenate([bx[:1, :, :2] * 0.9, bx[:1, :, :2] * 0.9 + 0.02 + 0.08 * bx[:1, :, 2:]], -1).astype(np.float32)
    s2 = np.abs(sc[2:3]) + np.float32(0.01)
    ns, nb, ni = core.nms_batch(torch.from_numpy(s2), torch.from_numpy(b2), 0.3, 150, 'min', assume_sorted=False, want_idx=True)
    o_s, o_b, o_i = O.nms(s2[0], b2[0], 0.3, 150, 'min')
    eq(ni[0], o_i.astype(np.int32), 'kept indices'); eq(ns[0], o_s, 'kept scores'); eq(nb[0], o_b, 'kept boxes')
    # (b) ron_eval on (nearly) every box of an image
    dec = O.flat_decode_anchors(O.anchors_all_layers(O.RON320))
    loc, pred, obj = synth.make_predictions(405, 1, 21250, 21, hot=300, dense=True)
    obj = np.maximum(obj, np.float32(0.5))
    boxes = O.decode(loc[0], dec)
    P = synth.split_layers(pred, LS); Ob = synth.split_layers(obj[..., None], LS); Bx = synth.split_layers(boxes[None], LS)
    old = ron_eval.FLAGS.select_threshold, ron_eval.FLAGS.objectness_thres
    try:
        ron_eval.FLAGS.select_threshold, ron_eval.FLAGS.objectness_thres = 1e-5, 0.03
        s, l, b = ron_eval.flaten_predict(P, Ob, [torch.from_numpy(t) for t in Bx])
        os_, ol, ob = O.flaten_predict([pred[0]], [obj[0]], [boxes], 0.03)
        assert os_.shape[0] > 16384
        eq(s, os_, 'scores')
        cs, cl, cb = ron_eval.tf_bboxes_nms_by_class(s, l, b, nms_threshold=0.45, keep_top_k=30, mode='min')
        rs, rl, rb = O.bboxes_nms_by_class(os_, ol, ob, 1e-5, 0.45, 30, 'min')
        eq(cs, rs, 'by-class scores'); eq(cl, rl, 'by-class labels'); eq(cb, rb, 'by-class boxes')
        cs, cl, cb = ron_eval.tf_bboxes_nms_by_class_v1(s, l, b, nms_threshold=0.45, keep_top_k=30, mode='min')
        rs, rl, rb = O.bboxes_nms_by_class_v1(os_, ol, ob, 1e-5, 21, 0.45, 30, 'min')
        eq(cs, rs, 'v1 scores'); eq(cl, rl, 'v1 labels'); eq(cb, rb, 'v1 boxes')
        cs, cl, cb = ron_eval.tf_bboxes_nms(s, l, b, nms_threshold=0.45, keep_top_k=100, mode='union')
        rs, rl, rb = O.bboxes_nms_agnostic(os_, ol, ob, 1e-5, 0.45, 100, 'union')
        eq(cs, rs, 'nms scores'); eq(cl, rl, 'nms labels'); eq(cb, rb, 'nms boxes')
    finally:
        ron_eval.FLAGS.select_threshold, ron_eval.FLAGS.objectness_thres = old
