#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the reference's OWN,
unmodified Python (/root/reference) over the NumPy emulation of TF-1 ops in
oracle/tf1_shim.  Run in the build container only (the GPU box has no
/root/reference); the .npz outputs are committed.

    python tests/golden/make_golden.py

What this pins: the reference's composition of ops (operand order, tie-breaking,
thresholds, padding) for anchors, joint encode, decode, select/clip/min-size/sort,
NMS ('min' and 'union'), TP/FP matching and AP.  Inputs come from
ron_tensorflow_b200.synth (seeded) plus hand-made edge cases; every fixture stores
the inputs it used, so tests never depend on generator stability.
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
assert 'numpy-shim' in tf.__version__
from nets import ron_vgg_320, ssd_vgg_512, ssd_vgg_300, ssd_common  # noqa: E402  (the reference)
import tf_extended as tfe  # noqa: E402  (the reference)
from ron_tensorflow_b200 import synth  # noqa: E402

T = tf.convert_to_tensor


def npy(x):
    return x.a if isinstance(x, tf.Tensor) else np.asarray(x)


def save(name, **arrs):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrs.items()})
    print('%-28s %8.1f KB  %s' % (name, os.path.getsize(path) / 1024., sorted(arrs.keys())))


# --------------------------------------------------------------------------- #
# anchors
# --------------------------------------------------------------------------- #
def gen_anchors():
    out = {}
    for tag, net in (('ron320', ron_vgg_320.RONNet()), ('ssd512', ssd_vgg_512.SSDNet()),
                     ('ssd300', ssd_vgg_300.SSDNet())):
        anchors = net.anchors(net.params.img_shape)
        for li, (y, x, h, w) in enumerate(anchors):
            out['%s_L%d_y' % (tag, li)] = y
            out['%s_L%d_x' % (tag, li)] = x
            out['%s_L%d_h' % (tag, li)] = h
            out['%s_L%d_w' % (tag, li)] = w
    save('anchors', **out)


# --------------------------------------------------------------------------- #
# joint encode (RONNet.bboxes_encode -> ssd_common.tf_ssd_bboxes_encode)
# --------------------------------------------------------------------------- #
def run_encode(net, anchors, labels, boxes, pos, ign):
    r = net.bboxes_encode(T(labels.astype(np.int64)), T(boxes.astype(np.float32)), anchors,
                          positive_threshold=pos, ignore_threshold=ign)
    lab = np.concatenate([npy(t).reshape(-1) for t in r[0]])
    loc = np.concatenate([npy(t).reshape(-1, 4) for t in r[1]])
    sco = np.concatenate([npy(t).reshape(-1) for t in r[2]])
    box = np.concatenate([npy(t).reshape(-1, 4) for t in r[3]])
    shapes = np.array([npy(t).shape[:3] for t in r[1]])
    return lab, loc, sco, box, shapes


def encode_cases():
    cases = {}
    # config 1: 5 synthetic GT boxes
    b, l = synth.make_gt(synth.image_seed(1, 0), 5)
    cases['cfg1_g5'] = (b, l, 0.5, 0.3)
    # trainer thresholds (ron_net.py:278), many GT
    b, l = synth.make_gt(synth.image_seed(2, 0), 50)
    cases['cfg2_g50_t056'] = (b, l, 0.56, 0.3)
    b, l = synth.make_gt(synth.image_seed(2, 1), 23)
    cases['cfg2_g23'] = (b, l, 0.5, 0.3)
    b, l = synth.make_gt(synth.image_seed(2, 2), 1)
    cases['cfg2_g1'] = (b, l, 0.56, 0.3)
    # grid-symmetric GT: exact float32 ties between anchors, two GT claiming one anchor,
    # the notebook GT (notebooks/ssd_tests.ipynb cell 14), a whole-image box
    b = np.array([[0.25, 0.25, 0.75, 0.75],
                  [0.25, 0.25, 0.75, 0.75],
                  [0.4, 0.4, 0.6, 0.6],
                  [0.0, 0.0, 1.0, 1.0],
                  [0.48, 0.136, 0.742, 0.552],
                  [0.024, 0.0227, 0.996, 0.997],
                  [0.1, 0.1, 0.2, 0.2],
                  [0.8, 0.8, 0.9, 0.9]], np.float32)
    l = np.array([3, 7, 12, 15, 12, 15, 1, 20], np.int64)
    cases['ties'] = (b, l, 0.5, 0.3)
    # all-zero overlap rows: a GT that only touches anchors outside the border mask and a
    # zero-area GT -> both are force-matched to anchor 0 (lowest GT index wins)
    b = np.array([[0.3, 0.3, 0.6, 0.7],
                  [0.5, 0.5, 0.5, 0.5],
                  [0.0, 0.0, 0.004, 0.004]], np.float32)
    l = np.array([5, 9, 2], np.int64)
    cases['zero_rows'] = (b, l, 0.5, 0.3)
    return cases


def gen_encode():
    net = ron_vgg_320.RONNet()
    anchors = net.anchors(net.params.img_shape)
    out = {}
    for name, (b, l, pos, ign) in encode_cases().items():
        lab, loc, sco, box, shapes = run_encode(net, anchors, l, b, pos, ign)
        out[name + '_in_boxes'] = b
        out[name + '_in_labels'] = l
        out[name + '_thr'] = np.array([pos, ign], np.float64)
        out[name + '_labels'] = lab
        out[name + '_loc'] = loc
        out[name + '_scores'] = sco
        out['anchor_boxes'] = box
        out['layer_shapes'] = shapes
    save('encode_ron320', **out)

    # do_dual_max_match with its two never-used flags, on a small random overlap matrix
    rng = np.random.Generator(np.random.PCG64(7))
    ov = rng.uniform(0, 1, size=(6, 40)).astype(np.float32)
    ov[ov < 0.35] = 0
    ov[2] = 0
    ov[:, 5] = ov[:, 6]
    mm = {'ov': ov}
    for ib in (True, False):
        for gf in (True, False):
            m, s = ssd_common.do_dual_max_match(T(ov), 0.5, 0.3, ignore_between=ib, gt_max_first=gf)
            mm['m_%d%d' % (ib, gf)] = npy(m)
            mm['s_%d%d' % (ib, gf)] = npy(s)
    save('dual_max_match', **mm)


# --------------------------------------------------------------------------- #
# post-process (eval_ron_network.py:223-252)
# --------------------------------------------------------------------------- #
def gen_postprocess():
    net = ron_vgg_320.RONNet()
    p = net.params
    anchors = net.anchors(p.img_shape)
    layer_sizes = [a[0].shape[0] * a[0].shape[1] * a[2].shape[0] for a in anchors]
    apc = [a[2].shape[0] for a in anchors]
    N = sum(layer_sizes)
    out = {}
    for tag, seed, hot, dense, K, M, thr in (('a', 31, 300, False, 400, 200, 0.45),
                                            ('dense', 32, 300, True, 200, 100, 0.4)):
        loc, pred, obj = synth.make_predictions(seed, 1, N, p.num_classes, hot=hot, dense=dense)
        loc_l = [T(t) for t in synth.split_layers(loc, layer_sizes, p.feat_shapes, apc)]
        pred_l = [T(t) for t in synth.split_layers(pred, layer_sizes, p.feat_shapes, apc)]
        obj_l = [T(t) for t in synth.split_layers(obj[..., None], layer_sizes, p.feat_shapes, apc)]
        # eval_ron_network.py:226-236
        dec = net.bboxes_decode(loc_l, anchors)
        filt = [tf.cast(tf.greater(o, 0.03), tf.float32) * pr for o, pr in zip(obj_l, pred_l)]
        rs, rb = net.detected_bboxes(filt, dec, select_threshold=0.01, nms_threshold=thr,
                                     clipping_bbox=[0., 0., 1., 1.], top_k=K, keep_top_k=M)
        classes = sorted(rs.keys())
        out[tag + '_seed'] = np.array([seed, hot, int(dense), K, M], np.int64)
        out[tag + '_nms_thr'] = np.array(thr, np.float64)
        out[tag + '_in_loc'] = loc
        out[tag + '_in_pred_sha'] = np.frombuffer(
            __import__('hashlib').sha256(pred.tobytes()).digest(), np.uint8)
        out[tag + '_in_obj_sha'] = np.frombuffer(
            __import__('hashlib').sha256(obj.tobytes()).digest(), np.uint8)
        out[tag + '_decoded'] = np.concatenate([npy(d).reshape(1, -1, 4) for d in dec], 1)
        out[tag + '_scores'] = np.stack([npy(rs[c]) for c in classes], 1)      # [1,C-1,M]
        out[tag + '_boxes'] = np.stack([npy(rb[c]) for c in classes], 1)       # [1,C-1,M,4]
        # the stage before NMS, for stage-wise parity
        s1, b1 = ssd_common.tf_ssd_bboxes_select(filt, dec, select_threshold=0.01,
                                                 num_classes=p.num_classes)
        b1 = tfe.bboxes_clip([0., 0., 1., 1.], b1)
        s1, b1 = net.bboxes_filter_min(s1, b1, K)
        s1, b1 = tfe.bboxes_sort(s1, b1, top_k=K)
        out[tag + '_topk_scores'] = np.stack([npy(s1[c]) for c in classes], 1)
        out[tag + '_topk_boxes'] = np.stack([npy(b1[c]) for c in classes], 1)
        if tag == 'a':
            # SSD order: select -> sort -> nms (ssd_vgg_512.py:182-201), no objectness
            ssd = ssd_vgg_512.SSDNet()
            rs2, rb2 = ssd.detected_bboxes(pred_l, dec, select_threshold=0.25,
                                           nms_threshold=0.45, top_k=100, keep_top_k=50)
            out['ssd_scores'] = np.stack([npy(rs2[c]) for c in classes], 1)
            out['ssd_boxes'] = np.stack([npy(rb2[c]) for c in classes], 1)
    save('postprocess_ron320', **out)


def gen_nms():
    rng = np.random.Generator(np.random.PCG64(11))
    out = {}
    K = 96
    c = rng.uniform(0.2, 0.8, size=(K, 2))
    s = rng.uniform(0.05, 0.4, size=(K, 2))
    boxes = np.concatenate([c - s / 2, c + s / 2], 1).astype(np.float32)
    scores = rng.uniform(0.02, 1, size=K).astype(np.float32)
    scores[10] = scores[40]                     # equal scores: lower index first
    scores[70:80] = 0                           # zero-padded slots are legal picks
    boxes[70:80] = 0
    boxes[20] = boxes[5]                        # duplicate box (overlap exactly 1)
    boxes[33, 2] = boxes[33, 0]                 # zero-area real box
    out['in_scores'] = scores
    out['in_boxes'] = boxes
    for mode in ('min', 'union'):
        for thr, M in ((0.45, 200), (0.3, 20), (0.7, 64)):
            rs, rb = tfe.bboxes_nms(T(scores), T(boxes), nms_threshold=thr, keep_top_k=M, mode=mode)
            out['%s_%g_%d_scores' % (mode, thr, M)] = npy(rs)
            out['%s_%g_%d_boxes' % (mode, thr, M)] = npy(rb)
    rs, rb = tfe.bboxes_nms_batch(T(np.stack([scores, scores[::-1].copy()])),
                                  T(np.stack([boxes, boxes[::-1].copy()])),
                                  nms_threshold=0.45, keep_top_k=32)
    out['batch_scores'] = npy(rs)
    out['batch_boxes'] = npy(rb)
    # bboxes_sort / bboxes_clip stand-alone
    ss, sb = tfe.bboxes_sort(T(scores[None]), T(boxes[None]), top_k=50)
    out['sort_scores'] = npy(ss)
    out['sort_boxes'] = npy(sb)
    wide = (boxes * 1.6 - 0.3).astype(np.float32)
    out['clip_in'] = wide
    out['clip_out'] = npy(tfe.bboxes_clip([0., 0., 1., 1.], T(wide)))
    save('nms', **out)


def gen_tpfp():
    rng = np.random.Generator(np.random.PCG64(13))
    B, M, Gm = 3, 24, 6
    gboxes, glabels, counts = synth.make_gt_batch(9, B, 2, Gm, num_classes=4, g_max=Gm)
    gdiff = (rng.uniform(size=(B, Gm)) < 0.25).astype(np.int64)
    out = {'gboxes': gboxes, 'glabels': glabels, 'gdiff': gdiff}
    d_s, d_b = {}, {}
    for c in (1, 2, 3):
        jit = rng.normal(0, 0.03, size=(B, M, 4)).astype(np.float32)
        src = rng.integers(0, Gm, size=(B, M))
        bx = np.take_along_axis(gboxes, src[..., None].repeat(4, -1), 1) + jit
        sc = np.sort(rng.uniform(0, 1, size=(B, M)).astype(np.float32), -1)[:, ::-1].copy()
        sc[:, -4:] = 0
        bx[:, -4:] = 0
        d_s[c], d_b[c] = sc, bx.astype(np.float32)
        out['det_scores_%d' % c] = d_s[c]
        out['det_boxes_%d' % c] = d_b[c]
    n, tp, fp, _ = tfe.bboxes_matching_batch(d_s.keys(), {c: T(v) for c, v in d_s.items()},
                                             {c: T(v) for c, v in d_b.items()},
                                             T(glabels), T(gboxes), T(gdiff), matching_threshold=0.5)
    for c in (1, 2, 3):
        out['n_gt_%d' % c] = npy(n[c])
        out['tp_%d' % c] = npy(tp[c])
        out['fp_%d' % c] = npy(fp[c])
        # metrics.py:169-175 filter, then precision_recall / AP (metrics.py:100-130,212-258)
        s = d_s[c].reshape(-1)
        t = npy(tp[c]).reshape(-1)
        f = npy(fp[c]).reshape(-1)
        mask = (t | f) & (s > np.float32(1e-4))
        s, t, f = s[mask], t[mask], f[mask]
        prec, rec = tfe.precision_recall(T(np.int64(npy(n[c]).sum())), s.shape[0], T(t), T(f), T(s))
        out['prec_%d' % c] = npy(prec)
        out['rec_%d' % c] = npy(rec)
        out['ap07_%d' % c] = npy(tfe.average_precision_voc07(prec, rec))
        out['ap12_%d' % c] = npy(tfe.average_precision_voc12(prec, rec))
    save('tpfp', **out)


# --------------------------------------------------------------------------- #
# ron_eval.py single-image variant.  ron_eval.py cannot be imported as a module (scipy.misc.imread,
# dataset / slim imports, flag definitions at import time), so the UNMODIFIED source of the four
# functions is cut out of the file with ``ast`` and executed over the same TF-1 shim.
# --------------------------------------------------------------------------- #
def gen_ron_eval():
    import ast
    src = open(os.path.join(REF, 'ron_eval.py')).read()
    want = {'flaten_predict', 'tf_bboxes_nms', 'filter_boxes', 'tf_bboxes_nms_by_class', 'tf_bboxes_nms_by_class_v1'}
    body = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name in want]

    class F(object):
        select_threshold, nms_threshold, objectness_thres, nms_topk, num_classes = 0.02, 0.4, 0.03, 20, 21
    ns = {'tf': tf, 'tfe': tfe, 'np': np, 'FLAGS': F}
    exec(compile(ast.Module(body=body, type_ignores=[]), 'ron_eval.py', 'exec'), ns)
    out = {}
    net = ron_vgg_320.RONNet()
    anchors = net.anchors(net.params.img_shape)
    LS, FS = [250, 1000, 4000, 16000], [(5, 5), (10, 10), (20, 20), (40, 40)]
    for tag, seed, hot, keep, mode, img in (('a', 91, 300, 20, 'union', (375, 500)), ('b', 92, 40, 200, 'min', (500, 333))):
        loc, pred, obj = synth.make_predictions(seed, 1, 21250, 21, hot=hot)
        P = [T(t) for t in synth.split_layers(pred, LS, FS, [10] * 4)]
        Ob = [T(t) for t in synth.split_layers(obj[..., None], LS, FS, [10] * 4)]
        Lc = [T(t) for t in synth.split_layers(loc, LS, FS, [10] * 4)]
        Bx = net.bboxes_decode(Lc, anchors)
        s, l, b = ns['flaten_predict'](P, Ob, Bx)
        out[tag + '_flat_scores'], out[tag + '_flat_labels'], out[tag + '_flat_boxes'] = npy(s), npy(l), npy(b)
        bbox_img = np.array([0., 0., 1., 1.], np.float32)
        b = tfe.bboxes.bboxes_clip(T(bbox_img), b)
        s, l, b = ns['filter_boxes'](s, l, b, 0.03, T(np.array(img, np.int32)), [320., 320.])
        out[tag + '_filt_scores'], out[tag + '_filt_labels'], out[tag + '_filt_boxes'] = npy(s), npy(l), npy(b)
        # the two per-class variants (ron_eval.py:212-366) on the same filtered boxes; keep_top_k 10 / 12
        # so that both the per-class cap and the final cut of _v1 (:361-363) are exercised
        cs, cl, cb = ns['tf_bboxes_nms_by_class'](s, l, b, nms_threshold=F.nms_threshold, keep_top_k=10, mode=mode)
        out[tag + '_bycls_scores'], out[tag + '_bycls_labels'], out[tag + '_bycls_boxes'] = npy(cs), npy(cl), npy(cb)
        cs, cl, cb = ns['tf_bboxes_nms_by_class_v1'](s, l, b, nms_threshold=F.nms_threshold, keep_top_k=12, mode=mode)
        out[tag + '_v1_scores'], out[tag + '_v1_labels'], out[tag + '_v1_boxes'] = npy(cs), npy(cl), npy(cb)
        s, l, b = ns['tf_bboxes_nms'](s, l, b, nms_threshold=F.nms_threshold, keep_top_k=keep, mode=mode)
        out[tag + '_nms_scores'], out[tag + '_nms_labels'], out[tag + '_nms_boxes'] = npy(s), npy(l), npy(b)
        ref = np.array([0.1, 0.05, 0.9, 0.95], np.float32)
        out[tag + '_resized'] = npy(tfe.bboxes.bboxes_resize(T(ref), b))
        out[tag + '_cfg'] = np.array([seed, hot, keep, mode == 'union', img[0], img[1]], np.int64)
        out[tag + '_in_pred_sha'] = np.frombuffer(__import__('hashlib').sha256(pred.tobytes()).digest(), np.uint8)
    save('ron_eval', **out)


# --------------------------------------------------------------------------- #
# mixed-class select + sort (nets/ssd_common.py:592-662, tf_extended/bboxes.py:27-57)
# --------------------------------------------------------------------------- #
def gen_mixed():
    out = {}
    LS, FS = [250, 1000, 4000, 16000], [(5, 5), (10, 10), (20, 20), (40, 40)]
    loc, pred, obj = synth.make_predictions(77, 1, 21250, 21, hot=300)
    boxes = np.clip(loc * np.float32(0.1) + np.float32(0.5), 0, 1).astype(np.float32)   # any "decoded" boxes
    P = [T(t) for t in synth.split_layers(pred, LS, FS, [10] * 4)]
    Bx = [T(t) for t in synth.split_layers(boxes, LS, FS, [10] * 4)]
    for tag, thr in (('none', None), ('thr', 0.05)):
        c, s, b = ssd_common.tf_ssd_bboxes_select_all_classes(P, Bx, select_threshold=thr)
        assert npy(c).dtype == np.int64 and np.array_equal(npy(b), boxes)
        out[tag + '_classes'], out[tag + '_scores'] = npy(c).astype(np.uint8), npy(s)      # classes < 21: stored as bytes
        c2, s2, b2 = tfe.bboxes.bboxes_sort_all_classes(c, s, b, top_k=300)
        out[tag + '_sorted_classes'], out[tag + '_sorted_scores'], out[tag + '_sorted_boxes'] = npy(c2), npy(s2), npy(b2)
    out['in_pred_sha'] = np.frombuffer(__import__('hashlib').sha256(pred.tobytes()).digest(), np.uint8)
    save('mixed_select', **out)


# --------------------------------------------------------------------------- #
# RON loss masks + smooth-L1 (nets/ron_vgg_320.py:635-771, nets/custom_layers.py:31-50)
# --------------------------------------------------------------------------- #
def gen_loss_masks():
    from nets import custom_layers
    out = {}
    FS = [(5, 5), (10, 10), (20, 20), (40, 40)]
    LS = [250, 1000, 4000, 16000]
    net = ron_vgg_320.RONNet()
    anchors = net.anchors(net.params.img_shape)
    for tag, seed, batch, g_lo, g_hi in (('a', 501, 2, 3, 12), ('b', 502, 1, 1, 1)):
        logits, loc, obj_logits, obj_pred = synth.make_loss_inputs(seed, batch)
        gb, gl, gc = synth.make_gt_batch(7, batch, g_lo, g_hi, first_image=seed)
        cls_l, loc_l, sc_l = [[] for _ in LS], [[] for _ in LS], [[] for _ in LS]
        for b in range(batch):
            c, l, s_, _ = net.bboxes_encode(T(gl[b, :gc[b]]), T(gb[b, :gc[b]]), anchors, positive_threshold=0.56, ignore_threshold=0.3)
            for i in range(len(LS)):
                cls_l[i].append(npy(c[i]).reshape(FS[i] + (10,)))
                loc_l[i].append(npy(l[i]))
                sc_l[i].append(npy(s_[i]).reshape(FS[i] + (10,)))
        gcls = [np.stack(v) for v in cls_l]
        gloc = [np.stack(v) for v in loc_l]
        gsc = [np.stack(v) for v in sc_l]
        split = lambda x, ch: [T(t) for t in synth.split_layers(x, LS, FS, [10] * 4)] if ch else None
        tf._rng = np.random.Generator(np.random.PCG64(seed + 1))
        del tf._stop_gradient_log[:], tf._loss_log[:], tf._random_log[:]
        ron_vgg_320.ron_losses([T(t) for t in synth.split_layers(logits, LS, FS, [10] * 4)],
                               [T(t) for t in synth.split_layers(loc, LS, FS, [10] * 4)],
                               [T(t) for t in synth.split_layers(obj_logits, LS, FS, [10] * 4)],
                               [T(t) for t in synth.split_layers(obj_pred[..., None], LS, FS, [10] * 4)],
                               [T(t) for t in gcls], [T(t) for t in gloc], [T(t) for t in gsc],
                               objness_threshold=0.03, negative_ratio=3.)
        sg = [npy(t) for t in tf._stop_gradient_log]
        # order of the tf.stop_gradient calls in ron_losses: final_neg_mask_objness :707, objness_pred_label :710,
        # cls_positive_mask :726, final_cls_neg_mask_objness :740, clipped gclasses :751, glocalisations :760,
        # cls_positive_mask again :764
        assert len(sg) == 7 and sg[0].dtype == bool and sg[1].dtype == np.int32 and sg[2].dtype == bool and sg[3].dtype == bool
        assert len(tf._random_log) == 2 and len(tf._loss_log) == 3
        out[tag + '_final_neg_mask_objness'] = np.packbits(sg[0])
        out[tag + '_objness_pred_label'] = np.packbits(sg[1].astype(bool))
        out[tag + '_cls_positive_mask'] = np.packbits(sg[2])
        out[tag + '_final_cls_neg_mask_objness'] = np.packbits(sg[3])
        out[tag + '_losses'] = np.array([float(npy(v)) for v in tf._loss_log], np.float32)     # cls, objectness, localisation
        # flat (layer-major) inputs of the mask stage, as ron_losses sees them (:660-675)
        flat = lambda ls: np.concatenate([np.asarray(t).reshape(-1) for t in ls])
        out[tag + '_gclasses'] = flat(gcls).astype(np.int8)
        out[tag + '_rand_sha'] = np.frombuffer(__import__('hashlib').sha256(tf._random_log[0].tobytes() + tf._random_log[1].tobytes()).digest(), np.uint8)
        out[tag + '_cfg'] = np.array([seed, batch, g_lo, g_hi], np.int64)
        # element-wise smooth-L1 of the first 64 rows of the flat order (bit-exact check)
        fl = np.concatenate([t.reshape(-1, 4) for t in synth.split_layers(loc, LS, FS, [10] * 4)])
        fg = np.concatenate([t.reshape(-1, 4) for t in gloc])
        pick = np.nonzero(flat(gcls) > 0)[0][:64]
        out[tag + '_sl1_pred'], out[tag + '_sl1_target'] = fl[pick], fg[pick]
        out[tag + '_sl1'] = npy(custom_layers.modified_smooth_l1(T(fl[pick]), T(fg[pick]), sigma=3.))
        out[tag + '_sl1_w'] = npy(custom_layers.modified_smooth_l1(T(fl[pick]), T(fg[pick]), 0.5, 2., sigma=1.))
    save('loss_masks', **out)


# --------------------------------------------------------------------------- #
# nets/np_methods.py: pure NumPy, so the UNMODIFIED reference functions run here (np.bool is restored for :232)
# --------------------------------------------------------------------------- #
def gen_np_methods():
    if not hasattr(np, 'bool'):
        np.bool = bool                      # removed in NumPy 1.24; np_methods.py:232 still spells it
    from nets import np_methods
    out = {}
    net = ssd_vgg_300.SSDNet()
    anchors = net.anchors((300, 300))
    LS = [38 * 38 * 4, 19 * 19 * 6, 10 * 10 * 6, 5 * 5 * 6, 3 * 3 * 4, 1 * 1 * 4]
    FS = [(38, 38), (19, 19), (10, 10), (5, 5), (3, 3), (1, 1)]
    AP = [4, 6, 6, 6, 4, 4]
    N = sum(LS)
    for tag, seed, thr, nms_thr in (('a', 611, 0.5, 0.45), ('b', 612, None, 0.3), ('c', 613, 0.05, 0.45)):
        loc, pred, _ = synth.make_predictions(seed, 1, N, 21, hot=120)
        P = synth.split_layers(pred, LS, FS, AP)
        Lc = synth.split_layers(loc, LS, FS, AP)
        # The reference's decode uses np.exp in float32, whose last bit depends on the NumPy build / CPU: its boxes
        # are kept for a tolerance check only, and the exact stages run (decode=False) on boxes from the oracle's
        # decode (correctly rounded exp), which every platform reproduces.
        if tag == 'a':                       # every 7th decoded row is enough for the tolerance check
            dec_ref = [np_methods.ssd_bboxes_decode(Lc[i], anchors[i]) for i in range(len(LS))]
            out[tag + '_decoded_7th'] = np.concatenate([d.reshape(-1, 4) for d in dec_ref]).astype(np.float32)[::7]
        from oracle import ron_oracle as _O
        flat = _O.decode(loc[0], _O.flat_decode_anchors(_O.anchors_all_layers(_O.SSD300)))
        dec = synth.split_layers(flat[None], LS, FS, AP)
        c, s_, b = np_methods.ssd_bboxes_select(P, dec, anchors, select_threshold=thr, img_shape=(300, 300), num_classes=21, decode=False)
        sha = lambda *a: np.frombuffer(__import__('hashlib').sha256(b''.join(np.ascontiguousarray(x).tobytes() for x in a)).digest(), np.uint8)
        # the selection can be long (15 k entries): its length and a digest of (classes int64, scores, boxes) are stored
        out[tag + '_sel_count'] = np.array([c.shape[0]], np.int64)
        out[tag + '_sel_sha'] = sha(c.astype(np.int64), s_.astype(np.float32), b.astype(np.float32))
        ref = np.array([0.05, 0.1, 0.9, 0.95], np.float32)
        b = np_methods.bboxes_clip(ref, b)
        out[tag + '_clipped_sha'] = sha(b.astype(np.float32))
        # np.argsort leaves the order of equal scores unspecified: the fixture must not depend on it
        assert np.array_equal(np.argsort(-s_)[:400], np.argsort(-s_, kind='stable')[:400]), 'tied scores inside the top-k'
        c, s_, b = np_methods.bboxes_sort(c, s_, b, top_k=400)
        out[tag + '_sort_classes'], out[tag + '_sort_scores'], out[tag + '_sort_boxes'] = c.astype(np.int16), s_, b
        c, s_, b = np_methods.bboxes_nms(c, s_, b, nms_threshold=nms_thr)
        out[tag + '_nms_classes'], out[tag + '_nms_scores'], out[tag + '_nms_boxes'] = c.astype(np.int16), s_, b
        out[tag + '_resized'] = np_methods.bboxes_resize(ref, b)
        out[tag + '_jaccard'] = np_methods.bboxes_jaccard(b[0], b)
        out[tag + '_intersection'] = np_methods.bboxes_intersection(b[0], b)
        out[tag + '_cfg'] = np.array([seed, -1 if thr is None else int(round(thr * 1000)), int(round(nms_thr * 1000))], np.int64)
        out[tag + '_in_pred_sha'] = np.frombuffer(__import__('hashlib').sha256(pred.tobytes() + loc.tobytes()).digest(), np.uint8)
    save('np_methods', **out)


# --------------------------------------------------------------------------- #
# datasets/voc_eval.py: result files + the PASCAL VOC evaluator (pure NumPy: the UNMODIFIED reference runs here)
# --------------------------------------------------------------------------- #
def gen_voc_eval():
    import tempfile, types, hashlib, contextlib, io
    if not hasattr(np, 'bool'):
        np.bool = bool                              # voc_eval.py:232
    sys.modules.setdefault('cv2', types.ModuleType('cv2'))     # imported at :16, never used
    from datasets import voc_eval

    class Dets(np.ndarray):
        """voc_eval.py:95 tests ``dets == []``, which NumPy 2 refuses to broadcast for a non-empty array."""
        def __eq__(self, other):
            return False if isinstance(other, list) else np.ndarray.__eq__(self, other)
        __hash__ = None

    out = {}
    seed = 4711
    ids, annots, all_boxes = synth.make_voc_eval_case(seed, 40)
    with tempfile.TemporaryDirectory() as tmp:
        synth.write_voc_tree(os.path.join(tmp, 'voc'), ids, annots)
        ev = voc_eval.DetectorEvalPascal(os.path.join(tmp, 'voc'), os.path.join(tmp, 'devkit'), 'test',
                                         output_dir=os.path.join(tmp, 'output_{}'))
        boxes = [[(b.view(Dets) if b.shape[0] else []) for b in per_cls] for per_cls in all_boxes]
        with contextlib.redirect_stdout(io.StringIO()):
            ev.write_voc_results_file(boxes)
            files = b''
            ap07, ap12, npt = [], [], []
            for ci, cls in enumerate(synth.VOC_CLASSES):
                fn = ev.get_voc_results_file_template(cls)
                txt = open(fn, 'rb').read()
                files += txt
                if ci == 14:
                    out['person_file'] = np.frombuffer(txt, np.uint8)
                cache = os.path.join(tmp, 'cache')
                rec, prec, a7 = ev.voc_eval(fn, cls, cache, ovthresh=0.5, use_07_metric=True)
                _, _, a12 = ev.voc_eval(fn, cls, cache, ovthresh=0.5, use_07_metric=False)
                ap07.append(a7); ap12.append(a12)
                npt.append(0 if np.isscalar(rec) else rec.shape[0])
                if ci in (6, 14):
                    out['rec_%d' % ci], out['prec_%d' % ci] = np.asarray(rec, np.float64), np.asarray(prec, np.float64)
    out['ap07'], out['ap12'] = np.array(ap07, np.float64), np.array(ap12, np.float64)
    out['n_dets'] = np.array(npt, np.int64)
    out['files_sha'] = np.frombuffer(hashlib.sha256(files).digest(), np.uint8)
    out['cfg'] = np.array([seed, 40], np.int64)
    save('voc_eval', **out)


# --------------------------------------------------------------------------- #
# datasets/pascalvoc_to_tfrecords.py: the reference's own converter writes a TFRecord file through the shim's
# tf.train.Example / tf.python_io.TFRecordWriter (public wire formats); committed as a fixture for the reader
# --------------------------------------------------------------------------- #
def gen_tfrecord():
    import tempfile, contextlib, io, shutil
    from datasets import pascalvoc_to_tfrecords as conv
    seed, n = 515, 12
    ids, annots, _ = synth.make_voc_eval_case(seed, n)
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, 'VOC2007') + os.sep
        synth.write_voc_tree(tmp, ids, annots, with_size=(375, 500, 3), with_images=True)
        os.makedirs(os.path.join(tmp, 'out'))
        with contextlib.redirect_stdout(io.StringIO()):
            conv.run(src, os.path.join(tmp, 'out'), name='voc_synth', shuffling=False)
        shutil.copy(os.path.join(tmp, 'out', 'voc_synth_000.tfrecord'), os.path.join(HERE, 'voc_synth_000.tfrecord'))
    print('%-30s %8.1f KB' % ('voc_synth_000.tfrecord', os.path.getsize(os.path.join(HERE, 'voc_synth_000.tfrecord')) / 1024.))


# --------------------------------------------------------------------------- #
# stand-alone pieces the fused path folds into kernels: RONNet.bboxes_filter_min (tensor and dict form),
# tfe.tensors.pad_axis, tfe.math.safe_divide
# --------------------------------------------------------------------------- #
def gen_filter_min():
    from tf_extended import tensors as tfe_tensors, math as tfe_math
    rng = np.random.Generator(np.random.PCG64(808))
    net = ron_vgg_320.RONNet()
    N = 300
    out = {}
    c = rng.uniform(0.1, 0.9, size=(3, N, 2))
    sz = rng.uniform(0.0, 0.09, size=(3, N, 2))               # about half of the boxes are below 0.03 in one side
    boxes = np.concatenate([c - sz / 2, c + sz / 2], -1).astype(np.float32)
    boxes[0, 7] = [0.2, 0.2, 0.23, 0.5]                        # width ok, height exactly 0.03 (float32): strict >
    boxes[1, :40] = 0                                          # zeroed (masked) entries never survive
    scores = rng.uniform(0, 1, size=(3, N)).astype(np.float32)
    out['in_scores'], out['in_boxes'] = scores, boxes
    for top_k in (50, 400):
        s, b = net.bboxes_filter_min(T(scores[0:1]), T(boxes[0:1]), top_k)
        out['k%d_scores' % top_k], out['k%d_boxes' % top_k] = npy(s), npy(b)
        ds, db = net.bboxes_filter_min({c + 1: T(scores[c:c + 1]) for c in range(3)}, {c + 1: T(boxes[c:c + 1]) for c in range(3)},
                                       top_k, minsize=0.04)
        for c in (1, 2, 3):
            out['k%d_dict_scores_%d' % (top_k, c)], out['k%d_dict_boxes_%d' % (top_k, c)] = npy(ds[c]), npy(db[c])
    x = rng.normal(size=(4, 5, 3)).astype(np.float32)
    out['pad_in'] = x
    for axis, size in ((0, 9), (1, 5), (1, 3), (2, 7)):
        out['pad_axis%d_size%d' % (axis, size)] = npy(tfe_tensors.pad_axis(T(x), 0, size, axis=axis))
    num = rng.normal(size=64).astype(np.float32)
    den = rng.normal(size=64).astype(np.float32)
    den[::7] = 0.
    out['div_num'], out['div_den'] = num, den
    out['div_out'] = npy(tfe_math.safe_divide(T(num), T(den), 'x'))
    save('filter_min', **out)


if __name__ == '__main__':
    if '--only-filter-min' in sys.argv:
        gen_filter_min()
        sys.exit(0)
    if '--only-tfrecord' in sys.argv:
        gen_tfrecord()
        sys.exit(0)
    if '--only-voc' in sys.argv:
        gen_voc_eval()
        sys.exit(0)
    if '--only-np' in sys.argv:
        gen_np_methods()
        sys.exit(0)
    if '--only-loss' in sys.argv:
        gen_loss_masks()
        sys.exit(0)
    if '--only-mixed' in sys.argv:
        gen_mixed()
        sys.exit(0)
    gen_ron_eval()
    gen_mixed()
    if '--only-ron-eval' in sys.argv:
        sys.exit(0)
    gen_anchors()
    gen_encode()
    gen_postprocess()
    gen_nms()
    gen_tpfp()
    gen_loss_masks()
    gen_np_methods()
    gen_voc_eval()
    gen_tfrecord()
    gen_filter_min()
