"""Drop-in for the hot-path half of the reference's ``nets/ssd_vgg_512.py``: ``SSDParams``,
``SSDNet.{anchors, bboxes_encode, bboxes_decode, detected_bboxes}``, ``ssd_anchor_one_layer``,
``ssd_anchors_all_layers`` (reference :48-101,150-201,286-358)."""
from collections import namedtuple

import numpy as np
import torch

from .. import core
from . import ssd_common
from .ron_vgg_320 import AnchorList, _nms_to_dicts, _fingerprint

# reference: nets/ssd_vgg_512.py:48-61
SSDParams = namedtuple('SSDParameters', ['img_shape', 'num_classes', 'no_annotation_label', 'feat_layers',
                                         'feat_shapes', 'anchor_size_bounds', 'anchor_sizes', 'anchor_ratios',
                                         'anchor_steps', 'anchor_offset', 'normalizations', 'prior_scaling'])


def ssd_anchor_one_layer(img_shape, feat_shape, sizes, ratios, step, offset=0.5, dtype=np.float32):
    """reference: nets/ssd_vgg_512.py:286-338."""
    a = core.AnchorSet('ssd', img_shape, [feat_shape], [sizes], [ratios], [step], offset, None)
    return tuple(v.astype(dtype) for v in a.as_reference_list()[0])


def ssd_anchors_all_layers(img_shape, layers_shape, anchor_sizes, anchor_ratios, anchor_steps, offset=0.5,
                           dtype=np.float32):
    """reference: nets/ssd_vgg_512.py:341-358."""
    a = core.AnchorSet('ssd', img_shape, layers_shape, anchor_sizes, anchor_ratios, anchor_steps, offset, None)
    out = AnchorList([tuple(v.astype(dtype) for v in t) for t in a.as_reference_list()])
    out.anchor_set = a
    out.fingerprint = _fingerprint(out)
    return out


class SSDNet(object):
    """reference: nets/ssd_vgg_512.py:63-201 (hot-path methods only)."""
    default_params = SSDParams(
        img_shape=(512, 512),
        num_classes=21,
        no_annotation_label=21,
        feat_layers=['block4', 'block7', 'block8', 'block9', 'block10', 'block11', 'block12'],
        feat_shapes=[(64, 64), (32, 32), (16, 16), (8, 8), (4, 4), (2, 2), (1, 1)],
        anchor_size_bounds=[0.10, 0.90],
        anchor_sizes=[(20.48, 51.2), (51.2, 133.12), (133.12, 215.04), (215.04, 296.96), (296.96, 378.88),
                      (378.88, 460.8), (460.8, 542.72)],
        anchor_ratios=[[2, .5], [2, .5, 3, 1. / 3], [2, .5, 3, 1. / 3], [2, .5, 3, 1. / 3], [2, .5, 3, 1. / 3],
                       [2, .5], [2, .5]],
        anchor_steps=[8, 16, 32, 64, 128, 256, 512],
        anchor_offset=0.5,
        normalizations=[20, -1, -1, -1, -1, -1, -1],
        prior_scaling=[0.1, 0.1, 0.2, 0.2])
    _clip_after_nms = False     # the reference commented the clip out (ssd_vgg_512.py:199-200)

    def __init__(self, params=None):
        self.params = params if isinstance(params, SSDParams) else type(self).default_params
        self._sets = {}

    def _set_for(self, img_shape):
        core._require_cuda()
        key = (tuple(img_shape), torch.cuda.current_device())
        if key not in self._sets:
            p = self.params
            self._sets[key] = core.AnchorSet('ssd', img_shape, p.feat_shapes, p.anchor_sizes, p.anchor_ratios,
                                             p.anchor_steps, p.anchor_offset, None)
        return self._sets[key]

    def _resolve(self, anchors):
        """See RONNet._resolve: the default handle only for None, the riding handle while the arrays are untouched,
        otherwise None (the caller's arrays are used as given)."""
        if anchors is None:
            return self._set_for(self.params.img_shape)
        return ssd_common.handle_of(anchors)

    def _resolve_set(self, anchors, allowed_borders=None):
        a = self._resolve(anchors)
        return a if a is not None else ssd_common.anchor_set_from_arrays(anchors, self.params.img_shape, allowed_borders)

    def anchors(self, img_shape, dtype=np.float32):
        """reference: nets/ssd_vgg_512.py:150-159."""
        a = self._set_for(img_shape)
        out = AnchorList([tuple(v.astype(dtype) for v in t) for t in a.as_reference_list()])
        out.anchor_set = a
        out.fingerprint = _fingerprint(out)
        return out

    def bboxes_encode(self, labels, bboxes, anchors, scope=None, positive_threshold=0.5, ignore_threshold=0.5,
                      allowed_borders=None):
        """reference: nets/ssd_vgg_512.py:161-171.  The reference wrapper still passes the
        pre-refactor argument list and raises TypeError; here the SSD
        anchors go through the same joint matcher with explicit borders (default: all inside)
        and the wrapper's ignore_threshold=0.5."""
        return ssd_common.tf_ssd_bboxes_encode(
            labels, bboxes, anchors, self.params.num_classes, self.params.img_shape, allowed_borders,
            self.params.no_annotation_label, positive_threshold=positive_threshold,
            ignore_threshold=ignore_threshold, prior_scaling=self.params.prior_scaling, scope=scope,
            _anchor_set=self._resolve(anchors))

    def bboxes_encode_batch(self, labels, bboxes, counts, anchors=None, positive_threshold=0.5,
                            ignore_threshold=0.5, want_matched=False, want_objness=False):
        return core.match_encode(self._resolve_set(anchors), bboxes, labels, counts, positive_threshold,
                                 ignore_threshold, self.params.prior_scaling, want_matched=want_matched,
                                 want_objness=want_objness)

    def bboxes_decode(self, feat_localizations, anchors, scope='ssd_bboxes_decode'):
        """reference: nets/ssd_vgg_512.py:173-180."""
        return ssd_common.tf_ssd_bboxes_decode(feat_localizations, anchors, prior_scaling=self.params.prior_scaling,
                                               scope=scope, _anchor_set=self._resolve(anchors))

    def detected_bboxes(self, predictions, localisations, select_threshold=None, nms_threshold=0.5,
                        clipping_bbox=None, top_k=400, keep_top_k=200):
        """reference: nets/ssd_vgg_512.py:182-201: select -> sort top_k -> NMS (no clip, no
        min-size); ``localisations`` are decoded boxes."""
        a = self._resolve(None)
        s, b, _ = core.decode_select_topk(a, localisations, predictions, None, 0.0, select_threshold, None, None,
                                          top_k, self.params.prior_scaling, loc_is_decoded=True)
        d_s, d_b = _nms_to_dicts(s, b, nms_threshold, keep_top_k)
        if self._clip_after_nms and clipping_bbox is not None:
            d_b = {c: core.clip(clipping_bbox, v) for c, v in d_b.items()}
        return d_s, d_b

    def detect(self, predictions, feat_localizations, select_threshold=None, nms_threshold=0.5, top_k=400,
               keep_top_k=200, mode='min'):
        """Fused SSD post-process from raw localisations (decode inside the select kernel)."""
        a = self._resolve(None)
        ns, nb, _ = core.select_nms(a, feat_localizations, predictions, None, 0.0, select_threshold, None, None, top_k,
                                    keep_top_k, nms_threshold, mode, self.params.prior_scaling)
        return ns, nb
