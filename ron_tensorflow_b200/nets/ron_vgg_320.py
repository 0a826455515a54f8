"""Drop-in for the hot-path half of the reference's ``nets/ron_vgg_320.py`` (lines 72-356):
``RONParams``, ``RONNet.{anchors, bboxes_encode, bboxes_decode, bboxes_filter_min,
detected_bboxes}``, ``ron_anchor_one_layer``, ``ron_anchors_all_layers`` -- same names,
argument order, defaults and return structure, with torch CUDA tensors where the reference
has TF tensors.  The network definition (reference :361-630) is out of scope; of the loss
(:635-771) the example masks and the localisation term are provided (``ron_loss_masks``,
``ron_localization_loss``: SURVEY.md section 8f rank 2), the two cross-entropy terms stay with the network.

Additions (never replacements): ``RONNet.bboxes_encode_batch`` (padded batch of images,
optional matched indices / objectness labels) and ``RONNet.detect`` (fused decode +
objectness gate + select + NMS straight from the network outputs).
"""
from collections import namedtuple

import numpy as np
import torch

from .. import core
from . import ssd_common

# reference: nets/ron_vgg_320.py:72-83
RONParams = namedtuple('SSDParameters', ['img_shape', 'num_classes', 'no_annotation_label', 'feat_layers',
                                         'feat_shapes', 'allowed_borders', 'anchor_sizes', 'anchor_ratios',
                                         'anchor_steps', 'anchor_offset', 'prior_scaling'])


class AnchorList(list):
    """The list of per-layer (y, x, h, w) NumPy arrays the reference returns, carrying the
    device-side anchor handle so bboxes_encode / bboxes_decode need no second upload.  ``fingerprint`` is a hash
    of the arrays at creation: a list whose arrays were edited afterwards no longer matches its handle and is treated
    as plain arrays."""
    anchor_set = None
    fingerprint = None


_fingerprint = ssd_common.fingerprint


def _anchor_set(kind, img_shape, layers_shape, anchor_sizes, anchor_ratios, anchor_steps, offset, borders):
    return core.AnchorSet(kind, img_shape, layers_shape, anchor_sizes, anchor_ratios, anchor_steps, offset, borders)


def ron_anchor_one_layer(img_shape, feat_shape, sizes, ratios, step, offset=0.5, dtype=np.float32):
    """reference: nets/ron_vgg_320.py:285-333.  Returns y[H,W,1], x[H,W,1], h[A], w[A]."""
    a = _anchor_set('ron', img_shape, [feat_shape], [sizes], [ratios], [step], offset, None)
    y, x, h, w = a.as_reference_list()[0]
    return y.astype(dtype), x.astype(dtype), h.astype(dtype), w.astype(dtype)


def ron_anchors_all_layers(img_shape, layers_shape, anchor_sizes, anchor_ratios, anchor_steps, offset=0.5,
                           dtype=np.float32, allowed_borders=None):
    """reference: nets/ron_vgg_320.py:336-355 (``allowed_borders`` is an addition: it lets the
    handle that rides on the returned list know the inside mask)."""
    a = _anchor_set('ron', img_shape, layers_shape, anchor_sizes, anchor_ratios, anchor_steps, offset,
                    allowed_borders)
    out = AnchorList([tuple(v.astype(dtype) for v in t) for t in a.as_reference_list()])
    out.anchor_set = a
    out.fingerprint = _fingerprint(out)
    return out


class RONNet(object):
    """reference: nets/ron_vgg_320.py:86-280 (hot-path methods only)."""
    # reference: nets/ron_vgg_320.py:97-124
    default_params = RONParams(
        img_shape=(320, 320),
        num_classes=21,
        no_annotation_label=21,
        feat_layers=['block7', 'block6', 'block5', 'block4'],
        feat_shapes=[(5, 5), (10, 10), (20, 20), (40, 40)],
        allowed_borders=[32, 16, 8, 4],
        anchor_sizes=[(224., 256.), (160., 192.), (96., 128.), (32., 64.)],
        anchor_ratios=[[1, 2, 3, 1. / 2, 1. / 3], [1, 2, 3, 1. / 2, 1. / 3],
                       [1, 2, 3, 1. / 2, 1. / 3], [1, 2, 3, 1. / 2, 1. / 3]],
        anchor_steps=[64, 32, 16, 8],
        anchor_offset=0.5,
        prior_scaling=[0.1, 0.1, 0.2, 0.2])

    def __init__(self, params=None):
        self.params = params if isinstance(params, RONParams) else RONNet.default_params
        self._sets = {}

    # ------------------------------------------------------------------ anchors
    def _set_for(self, img_shape):
        core._require_cuda()
        key = (tuple(img_shape), torch.cuda.current_device())
        if key not in self._sets:
            p = self.params
            self._sets[key] = _anchor_set('ron', img_shape, p.feat_shapes, p.anchor_sizes, p.anchor_ratios,
                                          p.anchor_steps, p.anchor_offset, p.allowed_borders)
        return self._sets[key]

    def _resolve(self, anchors):
        """The device handle to use for ``anchors``: the cached default set only when ``anchors`` is None; the handle
        riding on a list RONNet.anchors returned, as long as its arrays are untouched; None for any other list of
        (y, x, h, w) arrays -- the caller's values are then used as given (a content-hashed flat handle built by
        nets.ssd_common), like the reference, which always computes with the anchors it is handed."""
        if anchors is None:
            return self._set_for(self.params.img_shape)
        return ssd_common.handle_of(anchors)

    def _resolve_set(self, anchors, allowed_borders):
        """Like _resolve, but always a handle (for the batched forms that have no array path)."""
        a = self._resolve(anchors)
        return a if a is not None else ssd_common.anchor_set_from_arrays(anchors, self.params.img_shape, allowed_borders)

    def anchors(self, img_shape, dtype=np.float32):
        """reference: nets/ron_vgg_320.py:162-171."""
        a = self._set_for(img_shape)
        out = AnchorList([tuple(v.astype(dtype) for v in t) for t in a.as_reference_list()])
        out.anchor_set = a
        out.fingerprint = _fingerprint(out)
        return out

    # ------------------------------------------------------------------- encode
    def bboxes_encode(self, labels, bboxes, anchors, positive_threshold=0.5, ignore_threshold=0.3, scope=None):
        """reference: nets/ron_vgg_320.py:173-186 -> nets/ssd_common.py:337-414.  One image:
        labels [G] int64, bboxes [G,4].  Returns 4 lists over layers: labels flat [n_l] int64,
        localisations [H,W,A,4], scores flat [n_l], anchor corner boxes [H,W,A,4]."""
        return ssd_common.tf_ssd_bboxes_encode(
            labels, bboxes, anchors, self.params.num_classes, self.params.img_shape,
            self.params.allowed_borders, self.params.no_annotation_label,
            positive_threshold=positive_threshold, ignore_threshold=ignore_threshold,
            prior_scaling=self.params.prior_scaling, scope=scope, _anchor_set=self._resolve(anchors))

    def bboxes_encode_batch(self, labels, bboxes, counts, anchors=None, positive_threshold=0.5,
                            ignore_threshold=0.3, want_matched=False, want_objness=False):
        """Batched form: labels [B,Gmax] int64, bboxes [B,Gmax,4], counts [B] int32 ->
        dict(labels [B,N], loc [B,N,4], scores [B,N], matched?, objness?)."""
        return core.match_encode(self._resolve_set(anchors, self.params.allowed_borders), bboxes, labels, counts,
                                 positive_threshold, ignore_threshold, self.params.prior_scaling, want_matched=want_matched,
                                 want_objness=want_objness)

    def host_encoder(self, batch, g_max, anchors=None, slots=2, positive_threshold=0.5, ignore_threshold=0.3, **kw):
        """Batched encode for ground truth and targets in HOST memory (the reference encodes on the CPU inside its
        input pipeline): see core.HostEncoder -- submit(slot, boxes, labels, counts) / collect(slot)."""
        return core.HostEncoder(self._resolve_set(anchors, self.params.allowed_borders), batch, g_max, slots,
                                positive_threshold, ignore_threshold, self.params.prior_scaling, **kw)

    # ------------------------------------------------------------------- decode
    def bboxes_decode(self, feat_localizations, anchors, scope='ssd_bboxes_decode'):
        """reference: nets/ron_vgg_320.py:188-195."""
        return ssd_common.tf_ssd_bboxes_decode(feat_localizations, anchors, prior_scaling=self.params.prior_scaling,
                                               scope=scope, _anchor_set=self._resolve(anchors))

    def bboxes_filter_min(self, scores, bboxes, top_k, minsize=0.03, scope=None):
        """reference: nets/ron_vgg_320.py:196-233.  Keeps boxes with w > minsize and h > minsize
        (order preserving) and zero-pads to at least top_k.  The reference squeezes axis 0 and
        therefore only accepts batch 1 (:221); here every image of the batch is filtered and the
        result is padded to the longest row."""
        if isinstance(scores, dict) or isinstance(bboxes, dict):
            # all classes and images in two launches and one read-back.  Per class the reference pads to
            # max(its own survivors, top_k); rows are cut back to that width, so every class has the reference's shape
            keys = list(scores.keys())
            s, b = core.stack_classes(scores, keys), core.stack_classes(bboxes, keys)
            C, B = int(s.shape[0]), int(s.shape[1])
            os_, ob, counts = core.filter_min_rows(s.reshape(C * B, -1), b.reshape(C * B, -1, 4), top_k, minsize)
            os_, ob = os_.view(C, B, -1), ob.view(C, B, -1, 4)
            widths = [max(int(top_k), int(v)) for v in counts.reshape(C, B).max(1)]
            return ({c: os_[i, :, :widths[i]] for i, c in enumerate(keys)},
                    {c: ob[i, :, :widths[i]] for i, c in enumerate(keys)})
        s = core.as_cuda(scores, torch.float32)
        b = core.as_cuda(bboxes, torch.float32, s.device)
        os_, ob, _ = core.filter_min_rows(s, b, top_k, minsize)
        return os_, ob

    # ------------------------------------------------------------------ detect
    def detected_bboxes(self, predictions, localisations, select_threshold=None, nms_threshold=0.5,
                        clipping_bbox=None, top_k=400, keep_top_k=200):
        """reference: nets/ron_vgg_320.py:234-256.  ``localisations`` are DECODED boxes (the
        output of bboxes_decode), ``predictions`` the (already objectness-gated) class scores,
        both lists over layers.  select -> clip -> min-size 0.03 -> sort top_k -> NMS('min').
        Returns dicts class -> scores [B,keep_top_k], boxes [B,keep_top_k,4]."""
        a = self._resolve(None)
        s, b, _ = core.decode_select_topk(a, localisations, predictions, None, 0.0, select_threshold,
                                          clipping_bbox, 0.03, top_k, self.params.prior_scaling,
                                          loc_is_decoded=True)
        return _nms_to_dicts(s, b, nms_threshold, keep_top_k)

    def detect(self, predictions, feat_localizations, objness=None, objectness_threshold=0.03,
               select_threshold=None, nms_threshold=0.5, clipping_bbox=None, top_k=400, keep_top_k=200,
               minsize=0.03, mode='min', as_dict=False, want_idx=False):
        """Fused eval post-process of eval_ron_network.py:226-236: raw localisations + class
        scores (+ objectness) -> decode, objectness gate, select, clip, min-size, per-class top-k,
        NMS.  Returns scores [B,C-1,M], boxes [B,C-1,M,4] (and anchor indices when asked)."""
        a = self._resolve(None)
        ns, nb, aidx = core.select_nms(a, feat_localizations, predictions, objness, objectness_threshold, select_threshold,
                                       clipping_bbox, minsize, top_k, keep_top_k, nms_threshold, mode,
                                       self.params.prior_scaling, want_idx=want_idx)
        if as_dict:
            CM = ns.shape[1]
            return {c + 1: ns[:, c] for c in range(CM)}, {c + 1: nb[:, c] for c in range(CM)}
        if want_idx:
            return ns, nb, aidx
        return ns, nb


def _nms_to_dicts(s, b, nms_threshold, keep_top_k, mode='min'):
    B, CM, K = s.shape
    ns, nb, _ = core.nms_batch(s.view(B * CM, K), b.view(B * CM, K, 4), nms_threshold, keep_top_k, mode,
                               assume_sorted=True)
    ns, nb = ns.view(B, CM, -1), nb.view(B, CM, -1, 4)
    return {c + 1: ns[:, c] for c in range(CM)}, {c + 1: nb[:, c] for c in range(CM)}


# =========================================================================== #
# RON loss: example masks + localisation term (reference :635-771)
# =========================================================================== #
def ron_loss_masks(gclasses, objness_pred, rand_objness=None, rand_cls=None, objness_threshold=0.03,
                   negative_ratio=3., generator=None, localisations=None, glocalisations=None, beta=1. / 3):
    """reference nets/ron_vgg_320.py:686-740.  ``gclasses`` / ``objness_pred``: the encode labels and the
    objectness scores, flat or lists over layers (flattened and concatenated like :660-675).  The two
    ``tf.random_uniform`` draws of the reference (:705, :738) are ``rand_objness`` / ``rand_cls``; when
    omitted they are drawn on the device (``generator``: optional torch.Generator).
    Returns a dict: ``final_neg_mask_objness``, ``objness_pred_label`` (int32), ``cls_positive_mask``,
    ``final_cls_neg_mask_objness`` (bool tensors) and ``counts`` = float32 [n_positives, n_negtives,
    n_cls_positives, n_cls_negtives].  With ``localisations`` / ``glocalisations`` (flat [n,4] or lists over
    layers) the same launch also yields ``localization_loss`` (:760-764, see ron_localization_loss)."""
    def flat(x, dtype):
        if isinstance(x, (list, tuple)):
            return torch.cat([core.as_cuda(t, dtype).reshape(-1) for t in x], 0)
        return core.as_cuda(x, dtype).reshape(-1)
    g = flat(gclasses, torch.int64)
    o = flat(objness_pred, torch.float32)
    if rand_objness is None:
        rand_objness = torch.rand(g.shape, device=g.device, dtype=torch.float32, generator=generator)
    if rand_cls is None:
        rand_cls = torch.rand(g.shape, device=g.device, dtype=torch.float32, generator=generator)
    def flat4(x):
        if x is None:
            return None
        if isinstance(x, (list, tuple)):
            return torch.cat([core.as_cuda(t, torch.float32).reshape(-1, 4) for t in x], 0)
        return core.as_cuda(x, torch.float32).reshape(-1, 4)
    fl, fg = flat4(localisations), flat4(glocalisations)
    # the fused launch gives the VALUE of the localisation term; when the localisations take part in autograd the
    # term is computed by the differentiable form instead (one more launch), so loss.backward() reaches them
    need_grad = fl is not None and fl.requires_grad
    fo, lab, cp, fc, cnt, loss = core.loss_masks(g, o, flat(rand_objness, torch.float32), flat(rand_cls, torch.float32),
                                                 objness_threshold, negative_ratio, None if need_grad else fl,
                                                 None if need_grad else fg, 3., beta)
    out = dict(final_neg_mask_objness=fo, objness_pred_label=lab, cls_positive_mask=cp,
               final_cls_neg_mask_objness=fc, counts=cnt)
    if need_grad:
        out['localization_loss'] = core.localization_loss(fl, fg, cp, 3., beta)
    elif loss is not None:
        out['localization_loss'] = loss
    return out


def ron_localization_loss(localisations, glocalisations, cls_positive_mask, beta=1. / 3, sigma=3.):
    """reference nets/ron_vgg_320.py:760-764: beta * mean over the class positives of the row sums of
    modified_smooth_l1(localisations, glocalisations, sigma=3); 0 when there is no class positive."""
    def flat(x):
        if isinstance(x, (list, tuple)):
            return torch.cat([core.as_cuda(t, torch.float32).reshape(-1, 4) for t in x], 0)
        return core.as_cuda(x, torch.float32).reshape(-1, 4)
    return core.localization_loss(flat(localisations), flat(glocalisations), cls_positive_mask, sigma, beta)
