"""Drop-in for the post-process functions of the reference's ``ron_eval.py`` (single-image
evaluation): ``flaten_predict`` :111-144, ``filter_boxes`` :369-392,
the class-agnostic ``tf_bboxes_nms`` :146-210 and the per-class ``tf_bboxes_nms_by_class`` :212-291 /
``tf_bboxes_nms_by_class_v1`` :293-366 -- same names, argument order and defaults.  The
thresholds the reference reads from ``tf.app.flags.FLAGS`` live in the module-level ``FLAGS``
object below (same names and defaults, ron_eval.py:82-91).

All results have data-dependent shapes (``tf.boolean_mask``); every function returns CUDA
tensors trimmed to their real length (one device->host read of the count per mask).
"""
import numpy as np
import torch

from . import core


class _Flags(object):
    select_threshold = 0.6       # ron_eval.py:82-83
    nms_threshold = 0.4          # :84-85
    objectness_thres = 0.95      # :86-87
    nms_topk_percls = 10         # :88-89
    nms_topk = 20                # :90-91
    num_classes = 21             # :61-62


FLAGS = _Flags()


def flaten_predict(predictions, objness_pred, localisations):
    """reference ron_eval.py:111-144.  Lists over layers of [1,H,W,A,C] / [1,H,W,A,1] / [1,H,W,A,4]
    (``localisations`` are decoded boxes).  Returns scores [M,C] = objness * predictions, labels int64 [M]
    (arg-max class), boxes [M,4] for the anchors whose label is not background and whose objectness exceeds
    FLAGS.objectness_thres, in anchor order."""
    if int(predictions[0].shape[0]) > 1:
        raise ValueError('only batch_size 1 is supported.')
    scores, labels, mask = core.flaten_predict(predictions, objness_pred, FLAGS.objectness_thres)
    boxes = torch.cat([core.as_cuda(t, torch.float32, scores.device).reshape(-1, 4) for t in localisations], 0)
    idx = core.compact_indices(mask)
    return core.gather_rows(scores, idx), core.gather_rows(labels, idx), core.gather_rows(boxes, idx)


def filter_boxes(scores, labels, bboxes, min_size_ratio, image_shape, net_input_shape):
    """reference ron_eval.py:369-392: keep boxes with both sides > min_size and the centre inside the image."""
    # :374, float32 like the TF graph: max(0.0001, ratio * sqrt(float(h * w) / (net_h * net_w)))
    area = np.float32(int(image_shape[0]) * int(image_shape[1]))
    min_size = np.maximum(np.float32(0.0001),
                          np.float32(min_size_ratio) * np.sqrt(area / np.float32(net_input_shape[0] * net_input_shape[1])))
    b = core.as_cuda(bboxes, torch.float32).reshape(-1, 4)
    idx = core.compact_indices(core.filter_boxes_mask(b, float(min_size)))
    return (core.gather_rows(core.as_cuda(scores, torch.float32, b.device), idx),
            core.gather_rows(core.as_cuda(labels, torch.int64, b.device), idx), core.gather_rows(b, idx))


MAX_SORT = 1 << 24        # core.sort_topk: shared-memory sort up to 16 384 boxes, global-memory radix sort (ronk_sort_rows) above
MAX_KEEP = 2048           # ronk_nms_batch: the kept list lives in shared memory


def _check_limits(n, keep_top_k, what):
    """The reference's tf.nn.top_k(k = number of boxes) has no upper bound; the kernels behind these functions do."""
    if n >= MAX_SORT:
        raise ValueError('%s: %d boxes pass FLAGS.select_threshold / objectness_thres, the sort takes fewer than %d' % (what, n, MAX_SORT))
    if keep_top_k > MAX_KEEP:
        raise ValueError('%s: keep_top_k = %d, the NMS kernel keeps at most %d boxes per row' % (what, keep_top_k, MAX_KEEP))


def tf_bboxes_nms(scores, labels, bboxes, nms_threshold=0.5, keep_top_k=200, mode='union', scope=None):
    """reference ron_eval.py:146-210: class-agnostic greedy NMS on the best class score of every box.
    scores [M,C]; returns the kept (score [k], label [k], box [k,4]) in decreasing score order."""
    if mode not in ('union', 'min'):
        raise ValueError('unknown mode to use for nms.')
    s = core.as_cuda(scores, torch.float32)
    best, mask = core.rowmax_mask(s, FLAGS.select_threshold)                      # :149-151
    idx = core.compact_indices(mask)
    best = core.gather_rows(best, idx)
    labels = core.gather_rows(core.as_cuda(labels, torch.int64, s.device), idx)
    bboxes = core.gather_rows(core.as_cuda(bboxes, torch.float32, s.device).reshape(-1, 4), idx)
    n = int(best.shape[0])
    if n < 1:                                                                       # tf.cond(num_anchors < 1, ...) :210
        return best, labels, bboxes
    _check_limits(n, keep_top_k, 'tf_bboxes_nms')
    ss, sb, si = core.sort_topk(best.reshape(1, n), bboxes.reshape(1, n, 4), n, want_idx=True)    # :155-156
    ns, nb, ni = core.nms_batch(ss, sb, nms_threshold, keep_top_k, mode, assume_sorted=True, want_idx=True)
    kept = core.compact_indices((ni[0] >= 0).to(torch.uint8))
    pos = core.gather_rows(ni[0].contiguous(), kept)                # positions in the sorted list
    src = core.gather_rows(si[0].contiguous(), pos)                 # positions before sorting
    return core.gather_rows(ns[0].contiguous(), kept), core.gather_rows(labels, src), core.gather_rows(nb[0].contiguous(), kept)


def tf_bboxes_nms_by_class(scores, labels, bboxes, nms_threshold=0.5, keep_top_k=200, mode='min', scope=None):
    """reference ron_eval.py:212-291: one greedy NMS per class column over the boxes whose score in that
    column exceeds FLAGS.select_threshold (at most keep_top_k picks each); a box survives with the best of
    the scores it was kept for and that class as its new label.  scores [n,C]; returns (score [k], label
    int64 [k], box [k,4]) in the input order."""
    if mode not in ('union', 'min'):
        raise ValueError('unknown mode to use for nms.')
    if FLAGS.select_threshold < 0:
        raise ValueError('tf_bboxes_nms_by_class: a negative select_threshold is not supported')
    s = core.as_cuda(scores, torch.float32)
    b = core.as_cuda(bboxes, torch.float32, s.device).reshape(-1, 4)
    n = int(s.shape[0])
    if n < 1:                                                                       # tf.cond(num_anchors < 1, ...) :291
        return s, core.as_cuda(labels, torch.int64, s.device), b
    _check_limits(n, keep_top_k, 'tf_bboxes_nms_by_class')
    cs, cb = core.class_columns(s, b, FLAGS.select_threshold)                      # :228 + the transpose of :277
    ss, sb, si = core.sort_topk(cs, cb, n, want_idx=True)                          # :217-218
    ns, _, ni = core.nms_batch(ss, sb, nms_threshold, keep_top_k, mode, assume_sorted=True, want_idx=True)
    mx, lab, mask = core.keep_by_class(s, ni, ns, si, FLAGS.select_threshold)      # :263-264,282-288
    idx = core.compact_indices(mask)
    return core.gather_rows(mx, idx), core.gather_rows(lab, idx), core.gather_rows(b, idx)


def tf_bboxes_nms_by_class_v1(scores, labels, bboxes, nms_threshold=0.5, keep_top_k=200, mode='min', scope=None):
    """reference ron_eval.py:293-366: boxes above FLAGS.select_threshold (best class score) are sorted once;
    for every class 1..FLAGS.num_classes-1 a greedy NMS runs among the boxes carrying that label (at most
    keep_top_k picks each); the union of the survivors, cut to the first keep_top_k, is returned in
    decreasing score order."""
    if mode not in ('union', 'min'):
        raise ValueError('unknown mode to use for nms.')
    s = core.as_cuda(scores, torch.float32)
    best, mask = core.rowmax_mask(s, FLAGS.select_threshold)                      # :295-296
    idx = core.compact_indices(mask)
    best = core.gather_rows(best, idx)
    labels = core.gather_rows(core.as_cuda(labels, torch.int64, s.device), idx)
    bboxes = core.gather_rows(core.as_cuda(bboxes, torch.float32, s.device).reshape(-1, 4), idx)
    n = int(best.shape[0])
    if n < 1:                                                                       # :366
        return best, labels, bboxes
    _check_limits(n, keep_top_k, 'tf_bboxes_nms_by_class_v1')
    ss, sb, si = core.sort_topk(best.reshape(1, n), bboxes.reshape(1, n, 4), n, want_idx=True)    # :301-302
    sl = core.gather_rows(labels, si[0].contiguous())
    gs, gb, gp = core.group_by_label(sl, ss[0], sb[0], FLAGS.num_classes)          # nms_mask = (labels == cls) :340
    _, _, ni = core.nms_batch(gs, gb, nms_threshold, keep_top_k, mode, assume_sorted=True, want_idx=True)
    kept = core.compact_indices(core.mark_positions(ni, gp))[:keep_top_k]           # :344, :349-352
    return core.gather_rows(ss[0].contiguous(), kept), core.gather_rows(sl, kept), core.gather_rows(sb[0].contiguous(), kept)
