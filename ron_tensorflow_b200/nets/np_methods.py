"""Drop-in for the reference's ``nets/np_methods.py`` (the NumPy post-process of its notebooks and demo,
SURVEY.md section 8f rank 3): same names, argument order and defaults; every step is a libronk kernel and the
results are CUDA tensors (NumPy inputs are uploaded).

reference map: ssd_bboxes_decode :23-55, ssd_bboxes_select_layer :58-100, ssd_bboxes_select :103-131,
bboxes_sort :137-150, bboxes_clip :153-164, bboxes_resize :167-184, bboxes_jaccard :187-207,
bboxes_intersection :210-227, bboxes_nms :229-242.

Differences a caller can see: ``bboxes_sort`` orders equal scores by position (``np.argsort`` leaves their
order unspecified), and ``np.exp`` in the decode is replaced by the correctly rounded exponential every
other decode of this package uses (last-bit differences).
"""
import numpy as np
import torch

from .. import core

__all__ = ['ssd_bboxes_decode', 'ssd_bboxes_select_layer', 'ssd_bboxes_select', 'bboxes_sort', 'bboxes_clip',
           'bboxes_resize', 'bboxes_jaccard', 'bboxes_intersection', 'bboxes_nms']


def _anchor_set(anchor_bboxes):
    y, x, h, w = (np.asarray(v, np.float32) for v in anchor_bboxes)
    yy = np.broadcast_to(y.reshape(-1, 1), (y.size, h.size))
    xx = np.broadcast_to(x.reshape(-1, 1), (x.size, h.size))
    yxhw = np.stack([yy, xx, np.broadcast_to(h, yy.shape), np.broadcast_to(w, yy.shape)], -1).reshape(-1, 4)
    key = hash(yxhw.tobytes())
    cache = _anchor_set.__dict__.setdefault('cache', {})
    if key not in cache:
        if len(cache) >= 32:
            cache.clear()
        cache[key] = core.AnchorSet.flat((1, 1), yxhw, None)
    return cache[key]


def ssd_bboxes_decode(feat_localizations, anchor_bboxes, prior_scaling=[0.1, 0.1, 0.2, 0.2]):
    """reference :23-55.  feat_localizations [..., A, 4] for one layer, anchor_bboxes the layer's
    (y, x, h, w) tuple -> boxes (ymin, xmin, ymax, xmax), same shape."""
    loc = core.as_cuda(feat_localizations, torch.float32)
    aset = _anchor_set(anchor_bboxes)
    n = aset.N
    return core.decode(aset, loc.reshape(-1, n, 4), 0, prior_scaling).reshape(loc.shape)


def ssd_bboxes_select_layer(predictions_layer, localizations_layer, anchors_layer, select_threshold=0.5,
                            img_shape=(300, 300), num_classes=21, decode=True):
    """reference :58-100: classes, scores, bboxes of one layer."""
    if decode:
        localizations_layer = ssd_bboxes_decode(localizations_layer, anchors_layer)
    p = core.as_cuda(predictions_layer, torch.float32)
    p = p.reshape(-1, p.shape[-1])
    b = core.as_cuda(localizations_layer, torch.float32, p.device).reshape(-1, 4)
    return core.np_select(p, b, select_threshold)


def ssd_bboxes_select(predictions_net, localizations_net, anchors_net, select_threshold=0.5, img_shape=(300, 300),
                      num_classes=21, decode=True):
    """reference :103-131: the layers' selections concatenated."""
    parts = [ssd_bboxes_select_layer(predictions_net[i], localizations_net[i], anchors_net[i], select_threshold,
                                     img_shape, num_classes, decode) for i in range(len(predictions_net))]
    return tuple(torch.cat([p[k] for p in parts], 0) for k in range(3))


def bboxes_sort(classes, scores, bboxes, top_k=400):
    """reference :137-150: decreasing score, the first top_k."""
    s = core.as_cuda(scores, torch.float32)
    n = int(s.shape[0])
    k = min(int(top_k), n)
    if k < 1:
        return (core.as_cuda(classes, torch.int64, s.device)[:0], s[:0],
                core.as_cuda(bboxes, torch.float32, s.device).reshape(-1, 4)[:0])
    b = core.as_cuda(bboxes, torch.float32, s.device).reshape(1, n, 4)
    ss, sb, si = core.sort_topk(s.reshape(1, n), b, k, want_idx=True)
    cl = core.gather_i64(core.as_cuda(classes, torch.int64, s.device).reshape(1, n), si)
    return cl[0], ss[0], sb[0]


def bboxes_clip(bbox_ref, bboxes):
    """reference :153-164."""
    return core.np_clip(bbox_ref, bboxes)


def bboxes_resize(bbox_ref, bboxes):
    """reference :167-184."""
    return core.bboxes_resize(bbox_ref, bboxes)


def bboxes_jaccard(bboxes1, bboxes2):
    """reference :187-207 for one reference box against [n,4] boxes."""
    return core.overlap_ref(bboxes1, bboxes2, 'jaccard_np')


def bboxes_intersection(bboxes_ref, bboxes2):
    """reference :210-227 for one reference box against [n,4] boxes."""
    return core.overlap_ref(bboxes_ref, bboxes2, 'intersection_np')


def bboxes_nms(classes, scores, bboxes, nms_threshold=0.45):
    """reference :229-242: class-aware greedy NMS on score-sorted detections."""
    c = core.as_cuda(classes, torch.int64)
    s = core.as_cuda(scores, torch.float32, c.device)
    b = core.as_cuda(bboxes, torch.float32, c.device).reshape(-1, 4)
    idx = core.compact_indices(core.np_nms_keep(c, b, nms_threshold))
    return core.gather_rows(c, idx), core.gather_rows(s, idx), core.gather_rows(b, idx)
