"""Drop-in for the hot-path functions of the reference's ``tf_extended/bboxes.py``: same names,
argument order, defaults, dict-keyed-by-class inputs/outputs; torch CUDA tensors replace TF
tensors and every computation is a libronk kernel.

reference map: bboxes_sort_all_classes :27-57, bboxes_sort :60-101, bboxes_clip :105-144, bboxes_resize :147-171, bboxes_nms
:173-234, bboxes_nms_batch :262-302, bboxes_matching :316-404, bboxes_matching_batch :407-450,
bboxes_jaccard :527-554, bboxes_intersection :557-583.
"""
import torch

from .. import core

__all__ = ['bboxes_sort_all_classes', 'bboxes_sort', 'bboxes_clip', 'bboxes_resize', 'bboxes_nms', 'bboxes_nms_batch', 'bboxes_matching',
           'bboxes_matching_batch', 'bboxes_jaccard', 'bboxes_intersection']


def bboxes_sort_all_classes(classes, scores, bboxes, top_k=400, scope=None):
    """reference :27-57.  Mixed-class inputs: Batch x N classes / scores, Batch x N x 4 boxes ->
    the top_k by decreasing score (ties: lower index first) with their classes and boxes."""
    s, b, idx = core.sort_topk(scores, bboxes, top_k, want_idx=True)
    return core.gather_i64(classes, idx), s, b


def bboxes_sort(scores, bboxes, top_k=400, scope=None):
    """reference :60-101.  Batch x N scores / Batch x N x 4 boxes (or dicts of them) ->
    Batch x top_k, sorted by decreasing score, ties keep the lower index first."""
    if isinstance(scores, dict) or isinstance(bboxes, dict):
        # all classes in one launch: the kernel takes [S, N] rows, S = classes x images
        keys = list(scores.keys())
        s, b = core.stack_classes(scores, keys), core.stack_classes(bboxes, keys)
        C, B = int(s.shape[0]), int(s.shape[1])
        os_, ob, _ = core.sort_topk(s.reshape(C * B, -1), b.reshape(C * B, -1, 4), top_k)
        os_, ob = os_.view(C, B, -1), ob.view(C, B, -1, 4)
        return {c: os_[i] for i, c in enumerate(keys)}, {c: ob[i] for i, c in enumerate(keys)}
    s, b, _ = core.sort_topk(scores, bboxes, top_k)
    return s, b


def bboxes_clip(bbox_ref, bboxes, scope=None):
    """reference :105-144."""
    if isinstance(bboxes, dict):
        keys = list(bboxes.keys())
        shapes = {tuple(bboxes[c].shape) for c in keys}
        if len(shapes) == 1:                                   # one launch over all classes
            out = core.clip(bbox_ref, core.stack_classes(bboxes, keys))
            return {c: out[i] for i, c in enumerate(keys)}
        return {c: bboxes_clip(bbox_ref, bboxes[c]) for c in keys}
    return core.clip(bbox_ref, bboxes)


def bboxes_resize(bbox_ref, bboxes, name=None):
    """reference :147-171: translate by the reference corner, then divide by its size."""
    if isinstance(bboxes, dict):
        return {c: bboxes_resize(bbox_ref, bboxes[c]) for c in bboxes.keys()}
    return core.bboxes_resize(bbox_ref, bboxes)


def bboxes_nms(scores, bboxes, nms_threshold=0.5, keep_top_k=200, mode='min', scope=None):
    """reference :173-234.  One problem: scores [N], boxes [N,4] -> [max(keep_top_k, kept)]."""
    s = core.as_cuda(scores, torch.float32)
    b = core.as_cuda(bboxes, torch.float32, s.device)
    if s.numel() < 1:
        return s, b                                   # tf.cond(num_anchors < 1, ...) reference :234
    os_, ob, _ = core.nms_batch(s.reshape(1, -1), b.reshape(1, -1, 4), nms_threshold, keep_top_k, mode)
    return os_[0], ob[0]


def bboxes_nms_batch(scores, bboxes, nms_threshold=0.5, keep_top_k=200, scope=None, mode='min'):
    """reference :262-302 (always mode 'min' there; ``mode`` is an addition)."""
    if isinstance(scores, dict) or isinstance(bboxes, dict):
        # all classes in one launch (two with the re-sort): rows = classes x images
        keys = list(scores.keys())
        s, b = core.stack_classes(scores, keys), core.stack_classes(bboxes, keys)
        C, B = int(s.shape[0]), int(s.shape[1])
        os_, ob, _ = core.nms_batch(s.reshape(C * B, -1), b.reshape(C * B, -1, 4), nms_threshold, keep_top_k, mode)
        os_, ob = os_.view(C, B, -1), ob.view(C, B, -1, 4)
        return {c: os_[i] for i, c in enumerate(keys)}, {c: ob[i] for i, c in enumerate(keys)}
    os_, ob, _ = core.nms_batch(scores, bboxes, nms_threshold, keep_top_k, mode)
    return os_, ob


def bboxes_matching(label, scores, bboxes, glabels, gbboxes, gdifficults, matching_threshold=0.5, scope=None):
    """reference :316-404.  One image, one class: returns (n_gbboxes, tp [N] bool, fp [N] bool)."""
    s = core.as_cuda(scores, torch.float32).reshape(1, 1, -1)
    b = core.as_cuda(bboxes, torch.float32, s.device).reshape(1, 1, -1, 4)
    n, tp, fp = _match_class(int(label), s, b, glabels, gbboxes, gdifficults, matching_threshold, batched=False)
    return n[0], tp[0], fp[0]


def _match_class(label, s, b, glabels, gbboxes, gdifficults, thr, batched=True):
    gl = core.as_cuda(glabels, torch.int64, s.device)
    gb = core.as_cuda(gbboxes, torch.float32, s.device)
    gd = core.as_cuda(gdifficults, torch.int64, s.device)
    if not batched:
        gl, gb, gd = gl.reshape(1, -1), gb.reshape(1, -1, 4), gd.reshape(1, -1)
    # the kernel matches class index c against label c + 1: present the class as "class 1"
    gl1 = torch.where(gl == label, torch.ones_like(gl), torch.zeros_like(gl))
    n, tp, fp = core.tpfp_match(s, b, gl1, gb, gd, thr)
    return n[:, 0], tp[:, 0], fp[:, 0]


def bboxes_matching_batch(labels, scores, bboxes, glabels, gbboxes, gdifficults, matching_threshold=0.5,
                          scope=None):
    """reference :407-450.  Dicts class -> [B,N] / [B,N,4]; returns (d_n_gbboxes, d_tp, d_fp, scores)."""
    if isinstance(scores, dict) or isinstance(bboxes, dict):
        classes = list(labels)
        if classes and classes == list(range(1, len(classes) + 1)):
            # classes 1..C-1 at once: one launch over [B, C-1, M]
            s = torch.stack([core.as_cuda(scores[c], torch.float32) for c in classes], 1)
            b = torch.stack([core.as_cuda(bboxes[c], torch.float32) for c in classes], 1)
            n, tp, fp = core.tpfp_match(s, b, glabels, gbboxes, gdifficults, matching_threshold)
            return ({c: n[:, i] for i, c in enumerate(classes)}, {c: tp[:, i] for i, c in enumerate(classes)},
                    {c: fp[:, i] for i, c in enumerate(classes)}, scores)
        d_n, d_tp, d_fp = {}, {}, {}
        for c in classes:
            d_n[c], d_tp[c], d_fp[c], _ = bboxes_matching_batch(c, scores[c], bboxes[c], glabels, gbboxes,
                                                                gdifficults, matching_threshold)
        return d_n, d_tp, d_fp, scores
    s = core.as_cuda(scores, torch.float32)
    b = core.as_cuda(bboxes, torch.float32, s.device)
    n, tp, fp = _match_class(int(labels), s.unsqueeze(1), b.unsqueeze(1), glabels, gbboxes, gdifficults,
                             matching_threshold)
    return n, tp, fp, scores


def bboxes_jaccard(bbox_ref, bboxes, name=None):
    """reference :527-554."""
    return core.overlap_ref(bbox_ref, bboxes, 'jaccard')


def bboxes_intersection(bbox_ref, bboxes, name=None):
    """reference :557-583."""
    return core.overlap_ref(bbox_ref, bboxes, 'intersection')
