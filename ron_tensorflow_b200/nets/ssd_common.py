"""Drop-in for the hot-path functions of the reference's ``nets/ssd_common.py``: same names,
argument order and return structure; torch CUDA tensors replace TF tensors.

reference map: areas :27-29, intersection :30-42, iou_matrix :43-47, do_dual_max_match :49-75,
tf_ssd_bboxes_encode_layer :77-147, tf_ssd_bboxes_encode :337-414, tf_ssd_bboxes_decode_layer
:448-474, tf_ssd_bboxes_decode :477-498, tf_ssd_bboxes_select_layer :504-549,
tf_ssd_bboxes_select :552-589.
"""
import hashlib

import numpy as np
import torch

from .. import core

_PS = [0.1, 0.1, 0.2, 0.2]
_flat_cache = {}


def _flat_set(img_shape, yxhw, borders):
    """Anchor handle for caller-supplied flattened anchors, cached by content."""
    a = np.ascontiguousarray(np.asarray(yxhw, np.float32))
    b = None if borders is None else np.ascontiguousarray(np.asarray(borders, np.int32))
    key = (tuple(img_shape), a.shape, hashlib.sha1(a.tobytes()).hexdigest(),
           None if b is None else hashlib.sha1(b.tobytes()).hexdigest(), torch.cuda.current_device())
    s = _flat_cache.get(key)
    if s is None:
        if len(_flat_cache) > 8:
            _flat_cache.clear()
        s = core.AnchorSet.flat(img_shape, a, b)
        _flat_cache[key] = s
    return s


def _anchors_to_flat(anchors):
    """The NumPy part of tf_ssd_bboxes_encode (reference :371-388): per layer corners, then the
    re-derived (cy, cx, h, w), flattened (H, W, A) and concatenated over layers."""
    cols, sizes, shapes = [[], [], [], []], [], []
    for (yref, xref, href, wref) in anchors:
        yref, xref, href, wref = (np.asarray(v, np.float32) for v in (yref, xref, href, wref))
        ymin_ = yref - href / np.float32(2.)
        xmin_ = xref - wref / np.float32(2.)
        ymax_ = yref + href / np.float32(2.)
        xmax_ = xref + wref / np.float32(2.)
        shapes.append(ymin_.shape)
        sizes.append(int(np.prod(ymin_.shape)))
        for k, v in enumerate(((ymin_ + ymax_) / np.float32(2), (xmin_ + xmax_) / np.float32(2),
                               ymax_ - ymin_, xmax_ - xmin_)):
            cols[k].append(np.reshape(v, (-1)))
    flat = np.stack([np.concatenate(c) for c in cols], -1).astype(np.float32)
    return flat, sizes, shapes


def fingerprint(anchors):
    """Hash of a list of per-layer (y, x, h, w) arrays."""
    h = hashlib.sha1()
    for t in anchors:
        for v in t:
            a = np.ascontiguousarray(np.asarray(v))
            h.update(str(a.shape).encode() + str(a.dtype).encode() + a.tobytes())
    return h.hexdigest()


def handle_of(anchors):
    """The device handle riding on a list RONNet.anchors / SSDNet.anchors returned -- only while its arrays are the
    ones the handle was built from (the fingerprint taken at creation still matches); None otherwise."""
    a = getattr(anchors, 'anchor_set', None)
    if a is not None and getattr(anchors, 'fingerprint', None) == fingerprint(anchors):
        return a
    return None


def anchor_set_from_arrays(anchors, img_shape, allowed_borders):
    """Handle for a caller-supplied list of per-layer (y, x, h, w) arrays (reference :371-402), cached by content."""
    flat, sizes, _ = _anchors_to_flat(anchors)
    if allowed_borders is not None and len(allowed_borders) != len(sizes):
        raise IndexError('allowed_borders must have one entry per layer')
    borders = None if allowed_borders is None else \
        np.concatenate([np.full(n, b, np.int32) for n, b in zip(sizes, allowed_borders)])
    return _flat_set(img_shape, flat, borders)


# ----------------------------------------------------------------------------- IoU / matching
def areas(bboxes):
    """reference :27-29.  [G,4] -> [G,1]."""
    b = core.as_cuda(bboxes, torch.float32)
    return core.pairwise(b, None, 'areas')


def intersection(bboxes, gt_bboxes):
    """reference :30-42.  [G,4], [N,4] -> [G,N]."""
    return core.pairwise(core.as_cuda(bboxes, torch.float32), core.as_cuda(gt_bboxes, torch.float32), 'inter')


def iou_matrix(bboxes, gt_bboxes):
    """reference :43-47.  [G,4], [N,4] -> [G,N], 0 where the union is 0."""
    return core.pairwise(core.as_cuda(bboxes, torch.float32), core.as_cuda(gt_bboxes, torch.float32), 'iou')


def do_dual_max_match(overlap_matrix, high_thres, low_thres, ignore_between=True, gt_max_first=True):
    """reference :49-75.  overlap [G,N] -> (matched int64 [N], scores [N])."""
    return core.dual_max_match(core.as_cuda(overlap_matrix, torch.float32), high_thres, low_thres,
                               ignore_between, gt_max_first)


# ----------------------------------------------------------------------------- encode
def tf_ssd_bboxes_encode_layer(labels, bboxes, anchors_layer, num_classes, img_shape, allowed_border,
                               no_annotation_label, positive_threshold=0.5, ignore_threshold=0.3,
                               prior_scaling=_PS, dtype=torch.float32):
    """reference :77-147.  ``anchors_layer`` = (yref, xref, href, wref) broadcastable NumPy arrays,
    ``allowed_border`` a scalar or a per-anchor array.  Returns labels [N] int64, localisations
    [shape(anchors), 4], scores [N], anchor corner boxes [shape(anchors), 4]."""
    yref, xref, href, wref = (np.asarray(v, np.float32) for v in anchors_layer)
    shape = np.broadcast(yref, xref, href, wref).shape
    flat = np.stack([np.broadcast_to(v, shape).reshape(-1) for v in (yref, xref, href, wref)], -1)
    borders = np.broadcast_to(np.asarray(allowed_border), shape).reshape(-1) if np.ndim(allowed_border) \
        else np.full(flat.shape[0], int(allowed_border))
    aset = _flat_set(img_shape, flat, borders)
    r = _encode_one(aset, labels, bboxes, positive_threshold, ignore_threshold, prior_scaling)
    return (r['labels'][0], r['loc'][0].view(tuple(shape) + (4,)), r['scores'][0],
            aset.table(2).view(tuple(shape) + (4,)))


def _encode_one(aset, labels, bboxes, positive_threshold, ignore_threshold, prior_scaling):
    gl = core.as_cuda(labels, torch.int64, aset.device).reshape(1, -1)
    gb = core.as_cuda(bboxes, torch.float32, aset.device).reshape(1, -1, 4)
    if gl.shape[1] < 1:
        # tf.argmax over an empty GT axis fails in the reference; the trainer keeps >= 1 GT (ron_net.py:241-244)
        raise ValueError('bboxes_encode needs at least one ground-truth box')
    gc = torch.full((1,), gl.shape[1], dtype=torch.int32, device=aset.device)
    return core.match_encode(aset, gb, gl, gc, positive_threshold, ignore_threshold, prior_scaling)


def tf_ssd_bboxes_encode(labels, bboxes, anchors, num_classes, img_shape, allowed_borders, no_annotation_label,
                         positive_threshold=0.5, ignore_threshold=0.3, prior_scaling=_PS, dtype=torch.float32,
                         scope='ssd_bboxes_encode', _anchor_set=None):
    """reference :337-414.  Joint matching over the concatenated anchors of ALL layers.
    Returns four lists over layers: labels flat [n_l], localisations [H,W,A,4], scores flat [n_l],
    anchor corner boxes [H,W,A,4].  (The reference only works for exactly 4 layers, :385-388;
    any number of layers is accepted here with identical results for 4.)"""
    aset = _anchor_set if _anchor_set is not None else handle_of(anchors)
    if aset is not None and aset.gen_params is not None:
        aset = aset.with_borders(allowed_borders)
        sizes = aset.layer_sizes
        shapes = [(H, W, A) for (H, W, A, _) in aset.layers]
    else:
        flat, sizes, shapes = _anchors_to_flat(anchors)
        if allowed_borders is not None and len(allowed_borders) != len(sizes):
            raise IndexError('allowed_borders must have one entry per layer')
        borders = None if allowed_borders is None else \
            np.concatenate([np.full(n, b, np.int32) for n, b in zip(sizes, allowed_borders)])
        aset = _flat_set(img_shape, flat, borders)
    r = _encode_one(aset, labels, bboxes, positive_threshold, ignore_threshold, prior_scaling)
    lab, loc, sco, box = r['labels'][0], r['loc'][0], r['scores'][0], aset.table(2)
    ol, oc, os_, ob, o = [], [], [], [], 0
    for n, shp in zip(sizes, shapes):
        ol.append(lab[o:o + n])
        oc.append(loc[o:o + n].view(tuple(shp) + (4,)))
        os_.append(sco[o:o + n])
        ob.append(box[o:o + n].view(tuple(shp) + (4,)))
        o += n
    return ol, oc, os_, ob


# ----------------------------------------------------------------------------- decode
def tf_ssd_bboxes_decode_layer(feat_localizations, anchors_layer, prior_scaling=_PS):
    """reference :448-474.  feat_localizations [B,H,W,A,4] -> boxes [B,H,W,A,4]."""
    yref, xref, href, wref = (np.asarray(v, np.float32) for v in anchors_layer)
    t = core.as_cuda(feat_localizations, torch.float32)
    shape = tuple(t.shape[1:-1])
    flat = np.stack([np.broadcast_to(v, shape).reshape(-1) for v in (yref, xref, href, wref)], -1)
    aset = _flat_set((1, 1), flat, None)
    return core.decode(aset, t.reshape(t.shape[0], -1, 4), 0, prior_scaling).view(t.shape)


def tf_ssd_bboxes_decode(feat_localizations, anchors, prior_scaling=_PS, scope='ssd_bboxes_decode',
                         _anchor_set=None):
    """reference :477-498.  Lists over layers of [B,H,W,A,4]."""
    aset = _anchor_set if _anchor_set is not None else handle_of(anchors)
    if aset is None:
        return [tf_ssd_bboxes_decode_layer(feat_localizations[i], a, prior_scaling) for i, a in enumerate(anchors)]
    out = []
    for l, (H, W, A, o) in enumerate(aset.layers):
        t = core.as_cuda(feat_localizations[l], torch.float32, aset.device)
        out.append(core.decode(aset, t.reshape(t.shape[0], -1, 4), o, prior_scaling).view(t.shape))
    return out


# ----------------------------------------------------------------------------- select
def tf_ssd_bboxes_select_layer(predictions_layer, localizations_layer, select_threshold=None, num_classes=21,
                               ignore_class=0, scope=None):
    """reference :504-549.  Dicts class -> scores [B,n] (zero below the threshold), boxes [B,n,4]
    (zeroed likewise)."""
    p = core.as_cuda(predictions_layer, torch.float32)
    b = core.as_cuda(localizations_layer, torch.float32, p.device)
    p = p.reshape(p.shape[0], -1, p.shape[-1])
    b = b.reshape(b.shape[0], -1, b.shape[-1])
    s, bx = core.select_mask(p, b, select_threshold, ignore_class)
    classes = [c for c in range(num_classes) if c != ignore_class]
    return {c: s[:, i] for i, c in enumerate(classes)}, {c: bx[:, i] for i, c in enumerate(classes)}


def tf_ssd_bboxes_select(predictions_net, localizations_net, select_threshold=None, num_classes=21,
                         ignore_class=0, scope=None):
    """reference :552-589: per-layer select, concatenated over layers on axis 1."""
    l_s, l_b = [], []
    for i in range(len(predictions_net)):
        s, b = tf_ssd_bboxes_select_layer(predictions_net[i], localizations_net[i], select_threshold,
                                          num_classes, ignore_class)
        l_s.append(s)
        l_b.append(b)
    d_s, d_b = {}, {}
    for c in l_s[0].keys():
        d_s[c] = torch.cat([s[c] for s in l_s], dim=1)
        d_b[c] = torch.cat([b[c] for b in l_b], dim=1)
    return d_s, d_b


def tf_ssd_bboxes_select_layer_all_classes(predictions_layer, localizations_layer, select_threshold=None):
    """reference :592-628.  One class and score per anchor (mixed classes): classes int64 [B,n],
    scores [B,n], boxes [B,n,4] (the localisations are assumed decoded, :626-627)."""
    p = core.as_cuda(predictions_layer, torch.float32)
    b = core.as_cuda(localizations_layer, torch.float32, p.device)
    p = p.reshape(p.shape[0], -1, p.shape[-1])
    b = b.reshape(b.shape[0], -1, b.shape[-1])
    classes, scores = core.select_all_classes(p, select_threshold)
    return classes, scores, b


def tf_ssd_bboxes_select_all_classes(predictions_net, localizations_net, select_threshold=None, scope=None):
    """reference :631-662: per-layer select, concatenated over layers on axis 1."""
    l_c, l_s, l_b = [], [], []
    for i in range(len(predictions_net)):
        c, s, b = tf_ssd_bboxes_select_layer_all_classes(predictions_net[i], localizations_net[i], select_threshold)
        l_c.append(c)
        l_s.append(s)
        l_b.append(b)
    return torch.cat(l_c, dim=1), torch.cat(l_s, dim=1), torch.cat(l_b, dim=1)
