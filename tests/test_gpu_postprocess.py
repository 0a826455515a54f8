"""CUDA decode / select / top-k / NMS / TP-FP against the oracle and the golden vectors.
Bar: bit-exact kept-box indices, order, scores and boxes."""
import hashlib
import os

import numpy as np
import pytest

from oracle import ron_oracle as O
from ron_tensorflow_b200 import synth
from _util import need_cuda, eq

pytestmark = pytest.mark.gpu

LS = [250, 1000, 4000, 16000]
FS = [(5, 5), (10, 10), (20, 20), (40, 40)]
APC = [10, 10, 10, 10]


@pytest.fixture(scope='module')
def ron():
    need_cuda()
    from ron_tensorflow_b200.nets import ron_vgg_320
    net = ron_vgg_320.RONNet()
    return net, net.anchors(net.params.img_shape)


@pytest.fixture(scope='module')
def dec_anchors():
    return O.flat_decode_anchors(O.anchors_all_layers(O.RON320))


def _layers(x, shaped=True):
    return synth.split_layers(x, LS, FS if shaped else None, APC if shaped else None)


def _golden_inputs(g, tag):
    seed, hot, dense, K, M = [int(v) for v in g[tag + '_seed']]
    loc, pred, obj = synth.make_predictions(seed, 1, 21250, 21, hot=hot, dense=bool(dense))
    if hashlib.sha256(pred.tobytes()).digest() != g[tag + '_in_pred_sha'].tobytes():
        pytest.skip('numpy Generator stream differs from the one that made the fixture')
    return loc, pred, obj, K, M


@pytest.mark.parametrize('tag', ['a', 'dense'])
def test_detect_fused_golden(golden, ron, tag):
    net, anchors = ron
    g = golden('postprocess_ron320')
    loc, pred, obj, K, M = _golden_inputs(g, tag)
    from ron_tensorflow_b200 import core
    s, b, _ = core.decode_select_topk(anchors.anchor_set, _layers(loc), _layers(pred), _layers(obj[..., None]), 0.03,
                                      0.01, [0., 0., 1., 1.], 0.03, K)
    eq(s, g[tag + '_topk_scores'], 'top-k scores')
    eq(b, g[tag + '_topk_boxes'], 'top-k boxes')
    ns, nb = net.detect(_layers(pred), _layers(loc), _layers(obj[..., None]), 0.03, 0.01, float(g[tag + '_nms_thr']),
                        [0., 0., 1., 1.], K, M)
    eq(ns, g[tag + '_scores'], 'nms scores')
    eq(nb, g[tag + '_boxes'], 'nms boxes')


def test_reference_call_sequence_golden(golden, ron):
    """eval_ron_network.py:226-236 verbatim: bboxes_decode, objectness gate, detected_bboxes."""
    import torch
    net, anchors = ron
    g = golden('postprocess_ron320')
    loc, pred, obj, K, M = _golden_inputs(g, 'a')
    localisations = net.bboxes_decode([torch.as_tensor(t).cuda() for t in _layers(loc)], anchors)
    eq(torch.cat([t.reshape(1, -1, 4) for t in localisations], 1), g['a_decoded'], 'decoded')
    filtered = [(torch.as_tensor(o).cuda() > 0.03).float() * torch.as_tensor(p).cuda()
                for o, p in zip(_layers(obj[..., None]), _layers(pred))]
    rscores, rbboxes = net.detected_bboxes(filtered, localisations, select_threshold=0.01,
                                           nms_threshold=float(g['a_nms_thr']), clipping_bbox=[0., 0., 1., 1.],
                                           top_k=K, keep_top_k=M)
    assert sorted(rscores.keys()) == list(range(1, 21))
    eq(torch.stack([rscores[c] for c in range(1, 21)], 1), g['a_scores'], 'scores')
    eq(torch.stack([rbboxes[c] for c in range(1, 21)], 1), g['a_boxes'], 'boxes')


def test_ssd_order_golden(golden, ron):
    import torch
    from ron_tensorflow_b200.nets import ssd_vgg_512
    net, anchors = ron
    g = golden('postprocess_ron320')
    loc, pred, obj, K, M = _golden_inputs(g, 'a')
    dec = net.bboxes_decode(_layers(loc), anchors)
    ssd = ssd_vgg_512.SSDNet()
    ssd._sets[((512, 512), torch.cuda.current_device())] = anchors.anchor_set   # RON-shaped tensors, SSD pipeline
    rs, rb = ssd.detected_bboxes(_layers(pred), dec, select_threshold=0.25, nms_threshold=0.45, top_k=100,
                                 keep_top_k=50)
    eq(torch.stack([rs[c] for c in range(1, 21)], 1), g['ssd_scores'], 'scores')
    eq(torch.stack([rb[c] for c in range(1, 21)], 1), g['ssd_boxes'], 'boxes')


@pytest.mark.parametrize('dense,K,M,thr,mode', [(False, 400, 200, 0.45, 'min'), (False, 400, 200, 0.45, 'union'),
                                                (True, 400, 200, 0.45, 'min'), (False, 37, 11, 0.3, 'min')])
def test_detect_batch_vs_oracle(ron, dec_anchors, dense, K, M, thr, mode):
    net, anchors = ron
    B = 3
    loc, pred, obj = synth.make_predictions(77 + int(dense), B, 21250, 21, hot=300, dense=dense)
    ns, nb, ni = net.detect(_layers(pred, False), _layers(loc, False), _layers(obj, False), 0.03, 0.01, thr,
                            [0., 0., 1., 1.], K, M, mode=mode, want_idx=True)
    for b in range(B):
        o = O.detected_bboxes_image(pred[b], loc[b], dec_anchors, obj[b], 0.03, 0.01, thr, [0., 0., 1., 1.], K, M,
                                    mode=mode)
        eq(ni[b], o['idx'], 'kept anchor indices b=%d' % b)
        eq(ns[b], o['scores'], 'scores b=%d' % b)
        eq(nb[b], o['boxes'], 'boxes b=%d' % b)


def test_detect_81_classes(ron, dec_anchors):
    """config 5 shape: 81 classes (different shared-memory tile, 3 ballot words)."""
    net, anchors = ron
    loc, pred, obj = synth.make_predictions(99, 2, 21250, 81, hot=500)
    ns, nb, ni = net.detect(_layers(pred, False), _layers(loc, False), _layers(obj, False), 0.03, 0.005, 0.45,
                            [0., 0., 1., 1.], 200, 100, want_idx=True)
    for b in range(2):
        o = O.detected_bboxes_image(pred[b], loc[b], dec_anchors, obj[b], 0.03, 0.005, 0.45, [0., 0., 1., 1.], 200, 100)
        eq(ni[b], o['idx'], 'idx')
        eq(ns[b], o['scores'], 'scores')
        eq(nb[b], o['boxes'], 'boxes')


@pytest.mark.parametrize('mode', ['min', 'union'])
@pytest.mark.parametrize('thr,M', [(0.45, 200), (0.3, 20), (0.7, 64)])
def test_bboxes_nms_golden(golden, mode, thr, M):
    need_cuda()
    import ron_tensorflow_b200.tf_extended as tfe
    g = golden('nms')
    s, b = tfe.bboxes_nms(g['in_scores'], g['in_boxes'], nms_threshold=thr, keep_top_k=M, mode=mode)
    eq(s, g['%s_%g_%d_scores' % (mode, thr, M)], 'scores')
    eq(b, g['%s_%g_%d_boxes' % (mode, thr, M)], 'boxes')


@pytest.fixture(params=['auto', 'warp', 'wide', 'staged'])
def nms_variant(request, monkeypatch):
    """Every NMS kernel on every NMS case: the dispatch's own choice, one warp per segment (nms_kernel), a CTA per
    segment (nms_wide_kernel), and the two-stage CTA kernel for long lists (nms_staged_kernel; rows keeping <= 32 boxes
    fall back inside the library)."""
    env = {'auto': {}, 'warp': {'RONK_NMS_STAGED': '0', 'RONK_NMS_WIDE': '0'}, 'wide': {'RONK_NMS_STAGED': '0', 'RONK_NMS_WIDE': '4'},
           'staged': {'RONK_NMS_STAGED': '1'}}[request.param]
    for k in ('RONK_NMS_STAGED', 'RONK_NMS_WIDE'):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    return request.param


def test_nms_batch_sort_clip_golden(golden, nms_variant):
    need_cuda()
    import ron_tensorflow_b200.tf_extended as tfe
    g = golden('nms')
    sc = np.stack([g['in_scores'], g['in_scores'][::-1]])
    bx = np.stack([g['in_boxes'], g['in_boxes'][::-1]])
    s, b = tfe.bboxes_nms_batch(sc, bx, nms_threshold=0.45, keep_top_k=32)
    eq(s, g['batch_scores'], 'batch scores')
    eq(b, g['batch_boxes'], 'batch boxes')
    d_s, d_b = tfe.bboxes_nms_batch({3: sc, 7: sc[::-1].copy()}, {3: bx, 7: bx[::-1].copy()}, 0.45, 32)
    eq(d_s[3], g['batch_scores'], 'dict scores')
    eq(d_b[7], g['batch_boxes'][::-1], 'dict boxes')
    ss, sb = tfe.bboxes_sort(g['in_scores'][None], g['in_boxes'][None], top_k=50)
    eq(ss, g['sort_scores'], 'sort scores')
    eq(sb, g['sort_boxes'], 'sort boxes')
    eq(tfe.bboxes_clip([0., 0., 1., 1.], g['clip_in']), g['clip_out'], 'clip')
    with pytest.raises(ValueError):
        tfe.bboxes_sort(g['in_scores'][None], g['in_boxes'][None], top_k=500)    # tf.nn.top_k(k > N)
    with pytest.raises(ValueError):
        tfe.bboxes_nms(g['in_scores'], g['in_boxes'], mode='iou')              # bboxes.py:210


@pytest.mark.parametrize('K,M,mode', [(400, 200, 'min'), (400, 200, 'union'), (1000, 50, 'min'), (33, 40, 'min'),
                                      (2500, 300, 'union')])
def test_nms_random_vs_oracle(K, M, mode, nms_variant):
    need_cuda()
    from ron_tensorflow_b200 import core
    rng = np.random.Generator(np.random.PCG64(K * 7 + M))
    S = 6
    c = rng.uniform(0.1, 0.9, size=(S, K, 2))
    sz = np.exp(rng.uniform(np.log(0.02), np.log(0.5), size=(S, K, 2)))
    boxes = np.concatenate([c - sz / 2, c + sz / 2], -1).astype(np.float32)
    scores = rng.uniform(0.01, 1, size=(S, K)).astype(np.float32)
    tie = np.arange(7, K, 7)
    scores[:, tie] = scores[:, tie - 1]                                  # ties
    scores[0, K // 2:] = 0
    boxes[0, K // 2:] = 0
    s, b, ix = core.nms_batch(scores, boxes, 0.45, M, mode, want_idx=True)
    os_, ob, oi = O.nms_batch(scores, boxes, 0.45, M, mode)
    eq(ix, oi, 'kept positions')
    eq(s, os_, 'scores')
    eq(b, ob, 'boxes')


@pytest.mark.parametrize('case', ['outside_unit', 'zero_threshold', 'long_crowded', 'empty_boxes'])
def test_nms_special_cases_vs_oracle(case, nms_variant):
    """Cases that leave the fast path of the NMS kernels: coordinates outside [0, 1] (no saturating clamps), a threshold
    of 0 (a zero overlap suppresses), a long crowded list (6 000 candidates, most of them suppressed: the staged kernel's
    prefix filter and queue), boxes of zero area among the candidates."""
    need_cuda()
    from ron_tensorflow_b200 import core
    rng = np.random.Generator(np.random.PCG64({'outside_unit': 1, 'zero_threshold': 2, 'long_crowded': 3, 'empty_boxes': 4}[case]))
    S, K, M, thr, mode = 5, 700, 120, 0.45, 'min'
    if case == 'long_crowded':
        S, K, M = 3, 6000, 200
    c = rng.uniform(0.1, 0.9, size=(S, K, 2))
    sz = np.exp(rng.uniform(np.log(0.02), np.log(0.5), size=(S, K, 2)))
    boxes = np.concatenate([c - sz / 2, c + sz / 2], -1).astype(np.float32)
    scores = np.sort(rng.uniform(0.01, 1, size=(S, K)).astype(np.float32), axis=-1)[:, ::-1].copy()
    if case == 'outside_unit':
        boxes = boxes * np.float32(3.) - np.float32(0.7)
        mode = 'union'
        boxes[1] = (boxes[1] + np.float32(0.7)) / np.float32(3.)          # one row inside the unit square, 'union'
    if case == 'zero_threshold':
        thr = 0.
    if case == 'empty_boxes':
        boxes[:, ::5, 2] = boxes[:, ::5, 0]                               # zero height
        boxes[:, 3::11, 3] = boxes[:, 3::11, 1] - np.float32(0.01)        # negative width
    for sorted_ in (True, False):
        sc = scores if sorted_ else scores[:, rng.permutation(K)]
        s, b, ix = core.nms_batch(sc, boxes, thr, M, mode, assume_sorted=sorted_, want_idx=True)
        os_, ob, oi = O.nms_batch(sc, boxes, thr, M, mode)
        eq(ix, oi, 'kept positions'); eq(s, os_, 'scores'); eq(b, ob, 'boxes')


def test_tpfp_golden_and_random(golden):
    need_cuda()
    import torch
    import ron_tensorflow_b200.tf_extended as tfe
    g = golden('tpfp')
    d_s = {c: g['det_scores_%d' % c] for c in (1, 2, 3)}
    d_b = {c: g['det_boxes_%d' % c] for c in (1, 2, 3)}
    n, tp, fp, _ = tfe.bboxes_matching_batch(d_s.keys(), d_s, d_b, g['glabels'], g['gboxes'], g['gdiff'],
                                             matching_threshold=0.5)
    state = None
    for c in (1, 2, 3):
        eq(n[c], g['n_gt_%d' % c], 'n_gt')
        eq(tp[c], g['tp_%d' % c], 'tp')
        eq(fp[c], g['fp_%d' % c], 'fp')
    vals, state = tfe.streaming_tp_fp_arrays(n, tp, fp, {c: torch.as_tensor(v) for c, v in d_s.items()})
    for c in (1, 2, 3):
        prec, rec = tfe.precision_recall(*vals[c])
        eq(prec, g['prec_%d' % c], 'precision')
        eq(rec, g['rec_%d' % c], 'recall')
        assert abs(tfe.average_precision_voc07(prec, rec) - float(g['ap07_%d' % c])) < 1e-12
        assert abs(tfe.average_precision_voc12(prec, rec) - float(g['ap12_%d' % c])) < 1e-12
    # single-problem form (bboxes.py:316-404)
    n1, tp1, fp1 = tfe.bboxes_matching(2, d_s[2][1], d_b[2][1], g['glabels'][1], g['gboxes'][1], g['gdiff'][1])
    assert int(n1) == int(g['n_gt_2'][1])
    eq(tp1, g['tp_2'][1], 'single tp')
    eq(fp1, g['fp_2'][1], 'single fp')
    # random, 20 classes, 70 GT (more than one lane pass)
    from ron_tensorflow_b200 import core
    rng = np.random.Generator(np.random.PCG64(5))
    B, CM, M, G = 4, 20, 50, 70
    gb, gl, cnt = synth.make_gt_batch(8, B, 30, G, g_max=G)
    gd = (rng.uniform(size=(B, G)) < 0.2).astype(np.int64)
    src = rng.integers(0, G, size=(B, CM, M))
    det = np.take_along_axis(gb[:, None].repeat(CM, 1), src[..., None].repeat(4, -1), 2)
    det = (det + rng.normal(0, 0.02, size=det.shape)).astype(np.float32)
    sc = np.sort(rng.uniform(0, 1, size=(B, CM, M)).astype(np.float32), -1)[..., ::-1].copy()
    n, tp, fp = core.tpfp_match(sc, det, gl, gb, gd, 0.5)
    for b in range(B):
        for c in range(CM):
            on, otp, ofp = O.bboxes_matching(c + 1, sc[b, c], det[b, c], gl[b], gb[b], gd[b])
            assert int(n[b, c]) == on
            eq(tp[b, c], otp, 'tp b=%d c=%d' % (b, c))
            eq(fp[b, c], ofp, 'fp b=%d c=%d' % (b, c))


def test_fine_grained_functions(golden):
    """areas / intersection / iou_matrix / do_dual_max_match / select / jaccard one to one."""
    need_cuda()
    import torch
    from ron_tensorflow_b200.nets import ssd_common
    import ron_tensorflow_b200.tf_extended as tfe
    g = golden('dual_max_match')
    for ib in (1, 0):
        for gf in (1, 0):
            m, s = ssd_common.do_dual_max_match(g['ov'], 0.5, 0.3, ignore_between=bool(ib), gt_max_first=bool(gf))
            eq(m, g['m_%d%d' % (ib, gf)], 'matched %d%d' % (ib, gf))
            eq(s, g['s_%d%d' % (ib, gf)], 'scores %d%d' % (ib, gf))
    gt, _ = synth.make_gt(3, 9)
    _, cor, _ = O.encode_anchor_tables(O.anchors_all_layers(O.RON320), (320, 320), [32, 16, 8, 4])
    eq(ssd_common.iou_matrix(gt, cor), O.iou_matrix(gt, cor), 'iou_matrix')
    eq(ssd_common.areas(gt)[:, 0], O._areas(gt), 'areas')
    eq(tfe.bboxes_jaccard(gt[0], cor), O.bboxes_jaccard(gt[0], cor), 'jaccard')
    loc, pred, obj = synth.make_predictions(3, 2, 500, 21)
    d_s, d_b = ssd_common.tf_ssd_bboxes_select_layer(pred, loc, select_threshold=0.02)
    for c in (1, 7, 20):
        m = (pred[:, :, c] > np.float32(0.02)).astype(np.float32)
        eq(d_s[c], pred[:, :, c] * m, 'select scores')
        eq(d_b[c], loc * m[..., None], 'select boxes')
    assert 0 not in d_s and len(d_s) == 20


@pytest.mark.parametrize('path', ['sampled', 'unsampled', 'rebuild'])
@pytest.mark.parametrize('dense,K', [(False, 400), (True, 400), (False, 64), (True, 1500), (False, 1500), (True, 5000)])
def test_topk_paths_vs_oracle(ron, dec_anchors, path, dense, K):
    """The three ways the top-k kernel can arrive at its list -- sampled pivot + second scatter
    launch, one plain scatter launch, and the exact rebuild after a (forced) too-high pivot --
    must give the oracle's scores / anchor indices / boxes bit for bit."""
    from ron_tensorflow_b200 import core
    net, anchors = ron
    B = 2
    loc, pred, obj = synth.make_predictions(311 + int(dense) + K, B, 21250, 21, hot=300, dense=dense)
    s, bx, ix = core.decode_select_topk(anchors.anchor_set, _layers(loc, False), _layers(pred, False), _layers(obj, False),
                                        0.03, 0.01, [0., 0., 1., 1.], 0.03, K, want_idx=True,
                                        sampling=(path != 'unsampled'), _test_rebuild=(path == 'rebuild'))
    for b in range(B):
        boxes = O.decode(loc[b], dec_anchors)
        gate = (obj[b] > np.float32(0.03)).astype(np.float32)
        os_, ob, oi = O.select_topk_image(gate[:, None] * pred[b], boxes, 0.01, K, [0., 0., 1., 1.], 0.03)
        eq(ix[b], oi, 'anchor idx b=%d' % b)
        eq(s[b], os_, 'scores b=%d' % b)
        eq(bx[b], ob, 'boxes b=%d' % b)


@pytest.mark.parametrize('decimals', [2, 4])
def test_large_topk_with_score_ties(ron, dec_anchors, decimals):
    """top_k in the thousands with tied scores (the radix-sort select kernel sorts by the score bytes first and repairs
    runs of equal scores afterwards): scores rounded to 4 decimals give many short runs (ordered in place by the run's
    head thread), to 2 decimals runs of hundreds (the segment falls back to the full sort).  Order among equal scores =
    lower anchor first, like tf.nn.top_k."""
    from ron_tensorflow_b200 import core
    net, anchors = ron
    B, K = 2, 3000
    loc, pred, obj = synth.make_predictions(900 + decimals, B, 21250, 21, hot=300, dense=True)
    pred = np.round(pred, decimals).astype(np.float32)
    s, bx, ix = core.decode_select_topk(anchors.anchor_set, _layers(loc, False), _layers(pred, False), _layers(obj, False),
                                        0.03, 0.01, [0., 0., 1., 1.], 0.03, K, want_idx=True)
    ties = 0
    for b in range(B):
        boxes = O.decode(loc[b], dec_anchors)
        gate = (obj[b] > np.float32(0.03)).astype(np.float32)
        os_, ob, oi = O.select_topk_image(gate[:, None] * pred[b], boxes, 0.01, K, [0., 0., 1., 1.], 0.03)
        ties += int((np.diff(os_, axis=-1) == 0).sum())
        eq(ix[b], oi, 'anchor idx b=%d' % b)
        eq(s[b], os_, 'scores b=%d' % b)
        eq(bx[b], ob, 'boxes b=%d' % b)
    assert ties > 1000


def test_ssd512_postprocess_vs_oracle():
    """BASELINE config 4 shape on the post-process side: SSD-512 anchors (24 564, 7 layers, tiles
    that end mid-row), SSD order (no objectness gate, no clip, no min-size)."""
    need_cuda()
    from ron_tensorflow_b200.nets import ssd_vgg_512
    net = ssd_vgg_512.SSDNet()
    anchors = net.anchors(net.params.img_shape)
    ls = anchors.anchor_set.layer_sizes
    assert sum(ls) == 24564
    B = 2
    loc, pred, _ = synth.make_predictions(512, B, 24564, 21, hot=200)
    dec = O.flat_decode_anchors(O.anchors_all_layers(O.SSD512))
    ns, nb = net.detect(synth.split_layers(pred, ls), synth.split_layers(loc, ls), select_threshold=0.02,
                        nms_threshold=0.45, top_k=400, keep_top_k=200)
    for b in range(B):
        o = O.detected_bboxes_image(pred[b], loc[b], dec, None, None, 0.02, 0.45, None, 400, 200, min_size=None)
        eq(ns[b], o['scores'], 'scores b=%d' % b)
        eq(nb[b], o['boxes'], 'boxes b=%d' % b)


def test_crowded_81_classes_topk_10000(ron, dec_anchors):
    """BASELINE config 5 shape: 81 classes, dense scores (~10k candidates per class), K = 10 000
    (beyond the register/shared-memory paths of the top-k kernel) and K = 400."""
    from ron_tensorflow_b200 import core
    net, anchors = ron
    loc, pred, obj = synth.make_predictions(5005, 1, 21250, 81, hot=2000, dense=True)
    obj = np.maximum(obj, np.float32(0.05))                      # every anchor passes the objectness gate
    boxes = O.decode(loc[0], dec_anchors)
    for K in (10000, 400):
        s, bx, ix = core.decode_select_topk(anchors.anchor_set, _layers(loc, False), _layers(pred, False),
                                            _layers(obj, False), 0.03, 0.004, [0., 0., 1., 1.], 0.03, K, want_idx=True)
        os_, ob, oi = O.select_topk_image(pred[0], boxes, 0.004, K, [0., 0., 1., 1.], 0.03)
        assert int((os_ > 0).sum(1).max()) > 5000 or K == 400
        eq(ix[0], oi, 'anchor idx K=%d' % K)
        eq(s[0], os_, 'scores K=%d' % K)
        eq(bx[0], ob, 'boxes K=%d' % K)


def test_property_full_size_postprocess_batch256(ron, dec_anchors):
    """BASELINE config 3 at full size (batch 256): size-independent properties of the whole chain
    plus the oracle on a sample of the images.  Properties: scores sorted per (image, class); zero
    padding only at the tail; every kept pair respects the NMS threshold (min-area overlap < 0.45);
    running NMS again on the output keeps everything (idempotence); TP/FP flags are disjoint and
    TP count <= GT count."""
    import torch
    from ron_tensorflow_b200 import core
    net, anchors = ron
    B = 256
    loc, pred, obj = synth.make_predictions(3000, B, 21250, 21, hot=300)
    ns, nb, ni = net.detect(_layers(pred, False), _layers(loc, False), _layers(obj, False), 0.03, 0.01, 0.45,
                            [0., 0., 1., 1.], 400, 200, want_idx=True)
    assert ns.shape == (B, 20, 200)
    assert bool((ns[..., :-1] >= ns[..., 1:]).all()), 'scores must be sorted'
    real = ns > 0
    assert bool((real[..., :-1] | ~real[..., 1:]).all()), 'padding only at the tail'
    assert bool(((ni >= 0) == real).all())
    # pairwise min-area overlap among the kept boxes of a few (image, class) rows
    for b, c in [(0, 0), (17, 5), (255, 19), (128, 11)]:
        k = int(real[b, c].sum())
        bx = nb[b, c, :k].double()
        ih = (torch.minimum(bx[:, None, 2], bx[None, :, 2]) - torch.maximum(bx[:, None, 0], bx[None, :, 0])).clamp(min=0)
        iw = (torch.minimum(bx[:, None, 3], bx[None, :, 3]) - torch.maximum(bx[:, None, 1], bx[None, :, 1])).clamp(min=0)
        area = (bx[:, 2] - bx[:, 0]) * (bx[:, 3] - bx[:, 1])
        ov = ih * iw / torch.minimum(area[:, None], area[None, :])
        ov.fill_diagonal_(0)
        assert float(ov.max()) < 0.45 + 1e-6
    ns2, nb2, _ = core.nms_batch(ns.view(-1, 200), nb.view(-1, 200, 4), 0.45, 200, 'min', assume_sorted=True)
    assert torch.equal(ns2.view_as(ns), ns) and torch.equal(nb2.view_as(nb), nb), 'NMS must be idempotent'
    gb, gl, gc = synth.make_gt_batch(3, B, 1, 12, g_max=12)
    n_gt, tp, fp = core.tpfp_match(ns, nb, gl, gb, gl * 0, 0.5)
    assert not bool((tp & fp).any())
    assert bool((tp.sum(-1) <= n_gt).all())
    for b in (0, 101, 255):
        o = O.detected_bboxes_image(pred[b], loc[b], dec_anchors, obj[b], 0.03, 0.01, 0.45, [0., 0., 1., 1.], 400, 200)
        eq(ni[b], o['idx'], 'kept anchor indices b=%d' % b)
        eq(ns[b], o['scores'], 'scores b=%d' % b)
        eq(nb[b], o['boxes'], 'boxes b=%d' % b)


def test_cuda_graph_replay_matches_direct_calls(ron):
    """core.Graphed: the post-process chain and the encode captured into CUDA graphs give the same
    bits as direct calls, also after the inputs were refilled in place."""
    import torch
    from ron_tensorflow_b200 import core
    net, anchors = ron
    B = 4
    ins = []
    for seed in (11, 12):
        loc, pred, obj = synth.make_predictions(seed, B, 21250, 21, hot=300)
        ins.append([[torch.from_numpy(t).cuda() for t in _layers(x, False)] for x in (pred, loc, obj)])
    st = [[t.clone() for t in part] for part in ins[0]]
    fn = lambda p, l, o: net.detect(p, l, o, 0.03, 0.01, 0.45, [0., 0., 1., 1.], 400, 200)
    g = core.Graphed(fn, st[0], st[1], st[2])
    for k in (0, 1, 0):
        for part, src in zip(st, ins[k]):
            for d, s in zip(part, src):
                d.copy_(s)
        gs, gb = g.replay()
        ds, db = fn(*ins[k])
        assert torch.equal(gs, ds) and torch.equal(gb, db)
    boxes, labels, counts = synth.make_gt_batch(2, 16, 1, 50)
    d = [torch.from_numpy(x).cuda() for x in (boxes, labels, counts)]
    ge = core.Graphed(lambda: core.match_encode(anchors.anchor_set, d[0], d[1], d[2], 0.56, 0.3))
    r1 = ge.replay()
    r2 = core.match_encode(anchors.anchor_set, d[0], d[1], d[2], 0.56, 0.3)
    for k in ('labels', 'loc', 'scores'):
        assert torch.equal(r1[k], r2[k])
    r3 = ge.replay()
    assert torch.equal(r3['labels'], r2['labels'])


def test_device_tpfp_state_matches_host_accumulation():
    """tfe.TpFpDeviceState (records appended on the device, one copy at the end) against tfe.streaming_tp_fp_arrays
    (the reference's per-batch host accumulation, tf_extended/metrics.py:133-206): identical records in identical
    order for every class over several batches of different sizes, identical AP; zero scores and neither-TP-nor-FP
    entries are dropped; a too small buffer is reported."""
    need_cuda()
    import torch
    import ron_tensorflow_b200.tf_extended as tfe
    rng = np.random.Generator(np.random.PCG64(11))
    C = 7
    dev_state = tfe.TpFpDeviceState(C, capacity=1 << 16)
    host_state = None
    for B, M in ((3, 40), (5, 40), (1, 17), (4, 200)):
        sc = rng.uniform(0, 1, size=(B, C - 1, M)).astype(np.float32)
        sc[rng.uniform(size=sc.shape) < 0.15] = 0.
        sc[rng.uniform(size=sc.shape) < 0.05] = 1e-4                 # exactly at the filter's bound: dropped (strict >)
        tp = rng.uniform(size=sc.shape) < 0.3
        fp = (~tp) & (rng.uniform(size=sc.shape) < 0.6)
        ng = rng.integers(0, 5, size=(B, C - 1)).astype(np.int64)
        d = [torch.from_numpy(x).cuda() for x in (ng, tp, fp, sc)]
        dev_state.update(*d)
        _, host_state = tfe.streaming_tp_fp_arrays({c: ng[:, c - 1] for c in range(1, C)}, {c: tp[:, c - 1] for c in range(1, C)},
                                                   {c: fp[:, c - 1] for c in range(1, C)}, {c: sc[:, c - 1] for c in range(1, C)},
                                                   state=host_state)
    got = tfe.gather_tp_fp(dev_state, C)                              # no process group: the local state on the host
    for c in range(1, C):
        assert got[c].n_gt == host_state[c].n_gt
        eq(got[c].scores, host_state[c].scores, 'scores of class %d' % c)
        eq(got[c].tp, host_state[c].tp, 'tp')
        eq(got[c].fp, host_state[c].fp, 'fp')
        pa, ra = tfe.precision_recall(*got[c].value())
        pb, rb = tfe.precision_recall(*host_state[c].value())
        assert tfe.average_precision_voc07(pa, ra) == tfe.average_precision_voc07(pb, rb)
    assert int(dev_state.count_tensor().item()) == sum(host_state[c].scores.shape[0] for c in range(1, C))
    small = tfe.TpFpDeviceState(C, capacity=16)
    small.update(*d)
    with pytest.raises(RuntimeError):
        small.to_host()


def _pack_records(cls, scores, tp, fp):
    return ((scores.astype(np.float32).view(np.uint32).astype(np.int64) << 32) | (cls.astype(np.int64) << 8) |
            (fp.astype(np.int64) << 1) | tp.astype(np.int64))


def _oracle_ap(cls, scores, tp, fp, n_gt, C):
    out = {}
    for c in range(1, C):
        m = cls == c - 1
        prec, rec = O.precision_recall(int(n_gt[c - 1]), tp[m], fp[m], scores[m])
        out[c] = (prec, rec, O.average_precision_voc07(prec, rec), O.average_precision_voc12(prec, rec))
    return out


@pytest.mark.parametrize('case', ['random', 'ties', 'large', 'sparse_classes', 'many_classes', 'exact_tiles', 'one_record'])
def test_average_precision_records_vs_oracle(case):
    """tfe.average_precision_records (device: stable radix sort + per-class float64 curves) against the oracle's
    precision_recall / average_precision_voc07 / _voc12 (tf_extended/metrics.py:100-130, :212-258) class by class:
    precision and recall arrays bit-equal (exact integer sums, IEEE float64 quotients), VOC07 bit-equal, VOC12 to
    1e-12 (NumPy sums pairwise, the kernel tile by tile).  Cases: heavy score ties with mixed tp / fp (tf.nn.top_k's
    lower-index-first order decides), more records than one sort tile / AP tile, classes without records, classes
    without ground truth, padding entries, more than 255 classes (two class passes of the sort)."""
    need_cuda()
    import torch
    import ron_tensorflow_b200.tf_extended as tfe
    rng = np.random.Generator(np.random.PCG64({'random': 1, 'ties': 2, 'large': 3, 'sparse_classes': 4, 'many_classes': 5,
                                               'exact_tiles': 6, 'one_record': 7}[case]))
    C, n = {'random': (21, 5000), 'ties': (5, 6000), 'large': (4, 200000), 'sparse_classes': (12, 3000), 'many_classes': (301, 20000),
            'exact_tiles': (4, 16384 - 42), 'one_record': (3, 1)}[case]
    cls = rng.integers(0, C - 1, size=n)
    if case == 'exact_tiles':          # class sizes 4096 / 2048 / rest: boundaries on AP tiles; 16 384 entries with the padding: two full sort tiles
        cls = np.concatenate([np.zeros(4096, np.int64), np.ones(2048, np.int64), np.full(n - 6144, 2, np.int64)])[rng.permutation(n)]
    if case == 'sparse_classes':
        cls = rng.choice(np.array([0, 3, 4, 10]), size=n)                      # classes 2, 3, 6.. have no record at all
    scores = rng.uniform(1.1e-4, 1., size=n).astype(np.float32)
    if case == 'ties':
        scores = rng.choice(np.array([0.9, 0.5, 0.5000001, 0.25, 2e-4], np.float32), size=n)
    tp = rng.uniform(size=n) < 0.35
    fp = ~tp
    both = rng.uniform(size=n) < 0.02
    tp, fp = tp | both, fp | both                                              # (the matcher never sets both; the sums must not care)
    n_gt = rng.integers(1, max(2, n // (C - 1)), size=C - 1).astype(np.int64)
    if case in ('sparse_classes', 'random'):
        n_gt[0] = 0                                                            # recall = safe_div -> 0 everywhere
    rec = _pack_records(cls, scores, tp, fp)
    # padding entries in the middle and at the end (what the all-gather leaves behind a short rank)
    pad = np.full((37,), tfe.PAD_RECORD, np.int64) | (np.int64(0x3f000000) << 32)
    rec_in = np.concatenate([rec[:n // 3], pad, rec[n // 3:], pad[:5]])
    r = tfe.average_precision_records(torch.from_numpy(rec_in).cuda(), torch.from_numpy(n_gt).cuda(), C, curves=True)
    torch.cuda.synchronize()
    off = r['offsets'].cpu().numpy()
    assert off[0] == 0 and off[-1] == n
    want = _oracle_ap(cls, scores, tp, fp, n_gt, C)
    prec, recall = r['precision'].cpu().numpy(), r['recall'].cpu().numpy()
    srt = r['sorted'].cpu().numpy()
    ap07, ap12 = r['ap07'].cpu().numpy(), r['ap12'].cpu().numpy()
    for c in range(1, C):
        a, b = off[c - 1], off[c]
        m = cls == c - 1
        assert b - a == int(m.sum()), c
        o = O.topk_stable(scores[m], int(m.sum()))
        eq(srt[a:b], rec[m][o], 'sorted records of class %d' % c)
        eq(prec[a:b], want[c][0], 'precision of class %d' % c)
        eq(recall[a:b], want[c][1], 'recall of class %d' % c)
        assert ap07[c - 1] == want[c][2], (c, ap07[c - 1], want[c][2])
        assert abs(ap12[c - 1] - want[c][3]) <= 1e-12 * max(1., abs(want[c][3])), (c, ap12[c - 1], want[c][3])
    # no records at all
    z = tfe.average_precision_records(torch.zeros((0,), dtype=torch.int64, device='cuda'), torch.from_numpy(n_gt).cuda(), C)
    assert float(z['ap07'].abs().sum()) == 0. and float(z['ap12'].abs().sum()) == 0.


def test_device_state_average_precision_matches_host():
    """TpFpDeviceState.average_precision() (records never leave the device) equals precision_recall + AP of the host
    accumulators of tfe.streaming_tp_fp_arrays over several batches."""
    need_cuda()
    import torch
    import ron_tensorflow_b200.tf_extended as tfe
    rng = np.random.Generator(np.random.PCG64(12))
    C = 21
    dev_state = tfe.TpFpDeviceState(C, capacity=1 << 18)
    host_state = None
    for B, M in ((8, 200), (3, 200), (16, 200)):
        sc = np.round(rng.uniform(0, 1, size=(B, C - 1, M)), 2).astype(np.float32)      # two decimals: many equal scores
        sc[rng.uniform(size=sc.shape) < 0.3] = 0.
        tp = rng.uniform(size=sc.shape) < 0.2
        fp = (~tp) & (rng.uniform(size=sc.shape) < 0.7)
        ng = rng.integers(0, 9, size=(B, C - 1)).astype(np.int64)
        dev_state.update(*[torch.from_numpy(x).cuda() for x in (ng, tp, fp, sc)])
        _, host_state = tfe.streaming_tp_fp_arrays({c: ng[:, c - 1] for c in range(1, C)}, {c: tp[:, c - 1] for c in range(1, C)},
                                                   {c: fp[:, c - 1] for c in range(1, C)}, {c: sc[:, c - 1] for c in range(1, C)},
                                                   state=host_state)
    ap07, ap12, curves = dev_state.average_precision(curves=True)
    for c in range(1, C):
        p_, r_ = tfe.precision_recall(*host_state[c].value())
        eq(curves[c][0].cpu().numpy(), p_, 'precision of class %d' % c)
        eq(curves[c][1].cpu().numpy(), r_, 'recall of class %d' % c)
        assert ap07[c] == tfe.average_precision_voc07(p_, r_)
        assert abs(ap12[c] - tfe.average_precision_voc12(p_, r_)) <= 1e-12


def test_filter_min_pad_axis_safe_divide(golden):
    """RONNet.bboxes_filter_min stand-alone (tensor and dict forms, nets/ron_vgg_320.py:196-233), tfe.pad_axis and
    tfe.safe_divide against vectors recorded from the reference's own functions; then the batched form (every image of a
    batch, all classes in two launches) against the oracle image by image."""
    need_cuda()
    import torch
    import ron_tensorflow_b200.tf_extended as tfe
    from ron_tensorflow_b200.nets import ron_vgg_320
    net = ron_vgg_320.RONNet()
    g = golden('filter_min')
    for top_k in (50, 400):
        s, b = net.bboxes_filter_min(g['in_scores'][0:1], g['in_boxes'][0:1], top_k)
        eq(s, g['k%d_scores' % top_k], 'scores top_k=%d' % top_k)
        eq(b, g['k%d_boxes' % top_k], 'boxes')
        ds, db = net.bboxes_filter_min({c: g['in_scores'][c - 1:c] for c in (1, 2, 3)}, {c: g['in_boxes'][c - 1:c] for c in (1, 2, 3)},
                                       top_k, minsize=0.04)
        for c in (1, 2, 3):
            eq(ds[c], g['k%d_dict_scores_%d' % (top_k, c)], 'dict scores class %d top_k=%d' % (c, top_k))
            eq(db[c], g['k%d_dict_boxes_%d' % (top_k, c)], 'dict boxes')
    x = torch.from_numpy(g['pad_in']).cuda()
    for axis, size in ((0, 9), (1, 5), (1, 3), (2, 7)):
        eq(tfe.pad_axis(x, 0, size, axis=axis), g['pad_axis%d_size%d' % (axis, size)], 'pad_axis')
    eq(tfe.safe_divide(torch.from_numpy(g['div_num']).cuda(), torch.from_numpy(g['div_den']).cuda()), g['div_out'], 'safe_divide')
    # batched: B images x 4 classes, the reference semantics image by image (it only accepts batch 1, :221)
    rng = np.random.Generator(np.random.PCG64(4))
    B, N, K = 5, 700, 120
    c = rng.uniform(0.1, 0.9, size=(4, B, N, 2))
    sz = rng.uniform(0., 0.08, size=(4, B, N, 2))
    boxes = np.concatenate([c - sz / 2, c + sz / 2], -1).astype(np.float32)
    scores = rng.uniform(0, 1, size=(4, B, N)).astype(np.float32)
    ds, db = net.bboxes_filter_min({k + 1: scores[k] for k in range(4)}, {k + 1: boxes[k] for k in range(4)}, K)
    for k in range(4):
        rows = [O.bboxes_filter_min(scores[k, i], boxes[k, i], K) for i in range(B)]
        width = max(r[0].shape[0] for r in rows)
        assert tuple(ds[k + 1].shape) == (B, width)
        for i, (s_, b_) in enumerate(rows):
            eq(ds[k + 1][i, :s_.shape[0]], s_, 'class %d image %d' % (k + 1, i))
            eq(db[k + 1][i, :s_.shape[0]], b_, 'boxes')
            assert float(ds[k + 1][i, s_.shape[0]:].abs().sum()) == 0.


def test_ssd512_batch128_sampled():
    """BASELINE configs[3] at its stated batch size: SSD-512 anchor set, batch 128 -- post-process (SSD order) and
    match+encode, bit-exact against the oracle on a sample of the images."""
    need_cuda()
    from ron_tensorflow_b200.nets import ssd_vgg_512
    net = ssd_vgg_512.SSDNet()
    anchors = net.anchors(net.params.img_shape)
    ls = anchors.anchor_set.layer_sizes
    B = 128
    loc, pred, _ = synth.make_predictions(4000, B, 24564, 21, hot=300)
    oanch = O.anchors_all_layers(O.SSD512)
    dec = O.flat_decode_anchors(oanch)
    ns, nb = net.detect(synth.split_layers(pred, ls), synth.split_layers(loc, ls), select_threshold=0.01,
                        nms_threshold=0.45, top_k=400, keep_top_k=200)
    for b in (0, 63, 127):
        o = O.detected_bboxes_image(pred[b], loc[b], dec, None, None, 0.01, 0.45, None, 400, 200, min_size=None)
        eq(ns[b], o['scores'], 'scores b=%d' % b)
        eq(nb[b], o['boxes'], 'boxes b=%d' % b)
    del ns, nb
    boxes, labels, counts = synth.make_gt_batch(4, B, 1, 50)
    enc, cor, inside = O.encode_anchor_tables(oanch, O.SSD512.img_shape, None)
    import os
    for kernel in ('grid', 'generic'):
        os.environ['RONK_ENC_KERNEL'] = kernel
        try:
            r = net.bboxes_encode_batch(labels, boxes, counts, anchors, 0.5, 0.5, want_matched=True, want_objness=True)
        finally:
            del os.environ['RONK_ENC_KERNEL']
        for b in (0, 31, 77, 127):
            o = O.encode_image(labels[b, :counts[b]], boxes[b, :counts[b]], enc, cor, inside, 0.5, 0.5)
            eq(r['matched'][b], o['matched'].astype(np.int32), '%s matched[%d]' % (kernel, b))
            eq(r['labels'][b], o['labels'], 'labels')
            eq(r['scores'][b], o['scores'], 'scores')
            eq(r['loc'][b], o['loc'], 'loc')
            eq(r['objness'][b], o['objness'], 'objness')


@pytest.mark.parametrize('thr', [0.45, 0.02])
@pytest.mark.parametrize('tier_k', [1024, 1 << 30])
def test_crowded_detect_two_tier_topk_batch16(ron, dec_anchors, thr, tier_k, monkeypatch):
    """BASELINE config 5 shape through the fused detect path at batch 16: 81 classes, dense scores, top_k = 10 000, in
    one tier (the default) and in two (core.TIER_K = 1024: first 1024 candidates, then the full 10 000 for the rows that
    ran out before keep_top_k boxes were kept); thr = 0.02 suppresses almost everything, so many rows need the second
    tier."""
    from ron_tensorflow_b200 import core
    monkeypatch.setattr(core, 'TIER_K', tier_k)
    net, anchors = ron
    B, C, K, M = 16, 81, 10000, 200
    loc, pred, obj = synth.make_predictions(5005, B, 21250, C, hot=2000, dense=True)
    obj = np.maximum(obj, np.float32(0.05))
    ns, nb, ni = net.detect(_layers(pred, False), _layers(loc, False), _layers(obj, False), 0.03, 0.004, thr, [0., 0., 1., 1.],
                            K, M, want_idx=True)
    deep = 0
    for b in (0, 7, 15):
        o = O.detected_bboxes_image(pred[b], loc[b], dec_anchors, obj[b], 0.03, 0.004, thr, [0., 0., 1., 1.], K, M)
        eq(ni[b], o['idx'], 'kept anchor indices b=%d' % b)
        eq(ns[b], o['scores'], 'scores b=%d' % b)
        eq(nb[b], o['boxes'], 'boxes b=%d' % b)
        deep += int((o['nms_pos'].max(1) >= 1024).sum())
    if thr < 0.1:
        assert deep > 0, 'the second tier was not exercised'


def test_device_average_precision_over_nccl():
    """Two ranks with different record counts: NCCL gather that keeps the records on the device (padding marked) +
    device AP equals the host gather + NumPy AP on every rank (tools/nccl_ap_check.py under torchrun).  Needs 2 GPUs."""
    need_cuda()
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                        '--master-port', '29533', os.path.join(root, 'tools', 'nccl_ap_check.py')], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and 'nccl_ap_check OK' in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
