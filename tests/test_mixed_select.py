"""Mixed-class select / sort: tf_ssd_bboxes_select_all_classes
(nets/ssd_common.py:592-662) and bboxes_sort_all_classes (tf_extended/bboxes.py:27-57).  Golden vectors:
the reference's own functions over the TF-1 shim (tests/golden/make_golden.py --only-mixed)."""
import hashlib

import numpy as np
import pytest

from oracle import ron_oracle as O
from ron_tensorflow_b200 import synth
from _util import need_cuda, eq

LS = [250, 1000, 4000, 16000]
FS = [(5, 5), (10, 10), (20, 20), (40, 40)]


def _inputs(g):
    loc, pred, obj = synth.make_predictions(77, 1, 21250, 21, hot=300)
    if hashlib.sha256(pred.tobytes()).digest() != g['in_pred_sha'].tobytes():
        pytest.skip('numpy Generator stream differs from the one that made the fixture')
    boxes = np.clip(loc * np.float32(0.1) + np.float32(0.5), 0, 1).astype(np.float32)
    return pred, boxes


@pytest.mark.parametrize('tag,thr', [('none', None), ('thr', 0.05)])
def test_oracle_matches_reference(golden, tag, thr):
    g = golden('mixed_select')
    pred, boxes = _inputs(g)
    c, s = O.select_all_classes(pred, thr)
    assert c.dtype == np.int64 and np.array_equal(c, g[tag + '_classes'].astype(np.int64)) and np.array_equal(s, g[tag + '_scores'])
    c2, s2, b2 = O.sort_all_classes(c[0], s[0], boxes[0], 300)
    assert np.array_equal(c2, g[tag + '_sorted_classes'][0]) and np.array_equal(s2, g[tag + '_sorted_scores'][0])
    assert np.array_equal(b2, g[tag + '_sorted_boxes'][0])


@pytest.mark.gpu
@pytest.mark.parametrize('tag,thr', [('none', None), ('thr', 0.05)])
def test_cuda_matches_reference(golden, tag, thr):
    need_cuda()
    import torch
    from ron_tensorflow_b200.nets import ssd_common
    import ron_tensorflow_b200.tf_extended as tfe
    g = golden('mixed_select')
    pred, boxes = _inputs(g)
    P = synth.split_layers(pred, LS, FS, [10] * 4)
    Bx = synth.split_layers(boxes, LS, FS, [10] * 4)
    c, s, b = ssd_common.tf_ssd_bboxes_select_all_classes(P, Bx, select_threshold=thr)
    assert c.dtype == torch.int64
    eq(c, g[tag + '_classes'].astype(np.int64), 'classes'); eq(s, g[tag + '_scores'], 'scores'); eq(b, boxes, 'boxes')
    c2, s2, b2 = tfe.bboxes_sort_all_classes(c, s, b, top_k=300)
    eq(c2, g[tag + '_sorted_classes'], 'sorted classes'); eq(s2, g[tag + '_sorted_scores'], 'sorted scores')
    eq(b2, g[tag + '_sorted_boxes'], 'sorted boxes')


@pytest.mark.gpu
def test_cuda_vs_oracle_batch():
    need_cuda()
    from ron_tensorflow_b200.nets import ssd_common
    import ron_tensorflow_b200.tf_extended as tfe
    loc, pred, obj = synth.make_predictions(78, 3, 4999, 7, hot=50, dense=True)
    boxes = np.clip(loc * np.float32(0.1) + np.float32(0.5), 0, 1).astype(np.float32)
    for thr in (None, 0, 0.2):
        c, s, b = ssd_common.tf_ssd_bboxes_select_all_classes([pred], [boxes], select_threshold=thr)
        oc, os_ = O.select_all_classes(pred, thr)
        eq(c, oc, 'classes'); eq(s, os_, 'scores')
        c2, s2, b2 = tfe.bboxes_sort_all_classes(c, s, b, top_k=64)
        for i in range(3):
            r = O.sort_all_classes(oc[i], os_[i], boxes[i], 64)
            eq(c2[i], r[0], 'sorted classes'); eq(s2[i], r[1], 'sorted scores'); eq(b2[i], r[2], 'sorted boxes')
