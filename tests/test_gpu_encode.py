"""CUDA match+encode (through the reference-named API and the C ABI) against the oracle and the
golden vectors.  Bar: bit-exact labels / matched indices / scores; localisations bit-exact too
(both sides use correctly rounded exp/log), asserted <= 1e-5 relative as the north star states."""
import numpy as np
import pytest

from oracle import ron_oracle as O
from ron_tensorflow_b200 import synth
from _util import need_cuda, eq, close

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=['generic', 'grid'])
def encode_kernel(request, monkeypatch):
    """Every test of this module runs once per match+encode kernel: the batch size alone would send the small
    cases to the generic kernel only (csrc/match_encode.cu dispatches on B * N; RONK_ENC_KERNEL overrides it)."""
    monkeypatch.setenv('RONK_ENC_KERNEL', request.param)
    return request.param


@pytest.fixture(scope='module')
def ron():
    need_cuda()
    from ron_tensorflow_b200.nets import ron_vgg_320
    net = ron_vgg_320.RONNet()
    anchors = net.anchors(net.params.img_shape)
    return net, anchors


def _oracle_tables(spec):
    anchors = O.anchors_all_layers(spec)
    return anchors, O.encode_anchor_tables(anchors, spec.img_shape, spec.allowed_borders)


@pytest.mark.parametrize('which', ['ron320', 'ssd512', 'ssd300'])
def test_anchor_tables_bit_exact(golden, which):
    need_cuda()
    from ron_tensorflow_b200.nets import ron_vgg_320, ssd_vgg_512, ssd_vgg_300
    net, spec = {'ron320': (ron_vgg_320.RONNet(), O.RON320), 'ssd512': (ssd_vgg_512.SSDNet(), O.SSD512),
                 'ssd300': (ssd_vgg_300.SSDNet(), O.SSD300)}[which]
    anchors = net.anchors(net.params.img_shape)
    g = golden('anchors')
    for li, (y, x, h, w) in enumerate(anchors):
        for nm, v in (('y', y), ('x', x), ('h', h), ('w', w)):
            eq(v, g['%s_L%d_%s' % (which, li, nm)], '%s L%d %s' % (which, li, nm))
    aset = anchors.anchor_set
    oanch, (enc, cor, inside) = _oracle_tables(spec)
    eq(aset.table(0), O.flat_decode_anchors(oanch), 'decode anchors')
    eq(aset.table(1), enc, 'encode anchors')
    eq(aset.table(2), cor, 'corners')
    eq(aset.table(3).bool(), inside, 'inside mask')


@pytest.mark.parametrize('case', ['cfg1_g5', 'cfg2_g50_t056', 'cfg2_g23', 'cfg2_g1', 'ties', 'zero_rows'])
def test_bboxes_encode_golden(golden, ron, case):
    """RONNet.bboxes_encode (reference call shape, ron_net.py:277-278) vs vectors recorded from
    the reference's own code."""
    net, anchors = ron
    g = golden('encode_ron320')
    pos, ign = g[case + '_thr']
    lab, loc, sco, box = net.bboxes_encode(g[case + '_in_labels'], g[case + '_in_boxes'], anchors,
                                           positive_threshold=pos, ignore_threshold=ign)
    assert len(lab) == 4 and [tuple(t.shape) for t in loc] == [(5, 5, 10, 4), (10, 10, 10, 4), (20, 20, 10, 4),
                                                               (40, 40, 10, 4)]
    import torch
    eq(torch.cat([t.reshape(-1) for t in lab]), g[case + '_labels'], 'labels')
    eq(torch.cat([t.reshape(-1) for t in sco]), g[case + '_scores'], 'scores')
    eq(torch.cat([t.reshape(-1, 4) for t in loc]), g[case + '_loc'], 'loc')
    eq(torch.cat([t.reshape(-1, 4) for t in box]), g['anchor_boxes'], 'anchor boxes')


@pytest.mark.parametrize('pos,ign', [(0.56, 0.3), (0.5, 0.3)])
def test_encode_batch64_vs_oracle(ron, pos, ign):
    """BASELINE config 2: batch 64, 1-50 GT per image."""
    net, anchors = ron
    boxes, labels, counts = synth.make_gt_batch(2, 64, 1, 50)
    r = net.bboxes_encode_batch(labels, boxes, counts, anchors, pos, ign, want_matched=True, want_objness=True)
    _, (enc, cor, inside) = _oracle_tables(O.RON320)
    for b in range(64):
        o = O.encode_image(labels[b, :counts[b]], boxes[b, :counts[b]], enc, cor, inside, pos, ign)
        eq(r['matched'][b], o['matched'].astype(np.int32), 'matched[%d]' % b)
        eq(r['labels'][b], o['labels'], 'labels[%d]' % b)
        eq(r['scores'][b], o['scores'], 'scores[%d]' % b)
        eq(r['objness'][b], o['objness'], 'objness[%d]' % b)
        close(r['loc'][b], o['loc'], 1e-5, 'loc[%d]' % b)
        eq(r['loc'][b], o['loc'], 'loc[%d] bitwise' % b)
    # the workspace is left zeroed: a second call gives identical results
    r2 = net.bboxes_encode_batch(labels, boxes, counts, anchors, pos, ign, want_matched=True)
    eq(r2['matched'], r['matched'], 'second call matched')
    eq(r2['labels'], r['labels'], 'second call labels')


def test_encode_many_gt_and_flags(ron):
    """up to 200 GT boxes (config 5 shape) and the two never-used do_dual_max_match flags."""
    from ron_tensorflow_b200 import core
    net, anchors = ron
    boxes, labels, counts = synth.make_gt_batch(5, 6, 120, 200, num_classes=81)
    _, (enc, cor, inside) = _oracle_tables(O.RON320)
    for ib in (True, False):
        for gf in (True, False):
            r = core.match_encode(anchors.anchor_set, boxes, labels, counts, 0.5, 0.3, ignore_between=ib,
                                  gt_max_first=gf, want_matched=True)
            for b in range(6):
                o = O.encode_image(labels[b, :counts[b]], boxes[b, :counts[b]], enc, cor, inside, 0.5, 0.3,
                                   ignore_between=ib, gt_max_first=gf)
                eq(r['matched'][b], o['matched'].astype(np.int32), 'matched ib=%s gf=%s b=%d' % (ib, gf, b))
                eq(r['labels'][b], o['labels'], 'labels')
                eq(r['scores'][b], o['scores'], 'scores')
                eq(r['loc'][b], o['loc'], 'loc')


def test_encode_ssd512_anchor_set():
    """BASELINE config 4 shape (SSD-512 anchors, 24 564 anchors, 7 layers, all inside)."""
    need_cuda()
    from ron_tensorflow_b200.nets import ssd_vgg_512
    net = ssd_vgg_512.SSDNet()
    anchors = net.anchors(net.params.img_shape)
    boxes, labels, counts = synth.make_gt_batch(4, 8, 1, 50)
    r = net.bboxes_encode_batch(labels, boxes, counts, anchors, 0.5, 0.5, want_matched=True)
    _, (enc, cor, inside) = _oracle_tables(O.SSD512)
    assert inside.all() and enc.shape[0] == 24564
    for b in range(8):
        o = O.encode_image(labels[b, :counts[b]], boxes[b, :counts[b]], enc, cor, inside, 0.5, 0.5)
        eq(r['matched'][b], o['matched'].astype(np.int32), 'matched[%d]' % b)
        eq(r['labels'][b], o['labels'], 'labels')
        eq(r['scores'][b], o['scores'], 'scores')
        eq(r['loc'][b], o['loc'], 'loc')
    # list-per-layer form through the (repaired) SSD wrapper
    lab, loc, sco, box = net.bboxes_encode(labels[0, :counts[0]], boxes[0, :counts[0]], anchors)
    assert len(lab) == 7 and tuple(loc[0].shape) == (64, 64, 4, 4)


def test_encode_layer_with_plain_numpy_anchors(golden):
    """tf_ssd_bboxes_encode / _layer with anchors that carry no device handle (plain lists)."""
    need_cuda()
    from ron_tensorflow_b200.nets import ssd_common
    g = golden('encode_ron320')
    plain = [tuple(np.array(v) for v in t) for t in O.anchors_all_layers(O.RON320)]
    lab, loc, sco, box = ssd_common.tf_ssd_bboxes_encode(g['ties_in_labels'], g['ties_in_boxes'], plain, 21, (320, 320),
                                                         [32, 16, 8, 4], 21, 0.5, 0.3)
    import torch
    eq(torch.cat([t.reshape(-1) for t in lab]), g['ties_labels'], 'labels')
    eq(torch.cat([t.reshape(-1) for t in sco]), g['ties_scores'], 'scores')
    eq(torch.cat([t.reshape(-1, 4) for t in loc]), g['ties_loc'], 'loc')


def test_property_full_size_batch256(ron):
    """Size-independent properties at the bench size (batch 256): labels in range, positives
    carry the GT label of their matched index, every GT owns >= 1 anchor, unmatched loc is 0."""
    import torch
    net, anchors = ron
    boxes, labels, counts = synth.make_gt_batch(2, 256, 1, 50)
    r = net.bboxes_encode_batch(labels, boxes, counts, anchors, 0.56, 0.3, want_matched=True)
    m = r['matched'].long()
    lab = r['labels']
    gl = torch.as_tensor(labels, device=lab.device)
    assert int(lab.min()) >= -1 and int(lab.max()) <= 20
    pos = m >= 0
    assert torch.equal(lab[pos], torch.gather(gl, 1, m.clamp(min=0))[pos])
    assert torch.equal(lab[m == -2], torch.full_like(lab[m == -2], -1))
    assert bool((lab[m == -1] == 0).all())
    assert bool((r['loc'][~pos] == 0).all())
    cnt = torch.as_tensor(counts, device=lab.device)
    for b in range(0, 256, 17):
        owned = torch.unique(m[b][m[b] >= 0])
        assert owned.numel() == int(cnt[b]), 'image %d: every GT must be matched by at least one anchor' % b


def test_encode_batch256_vs_oracle_sampled(ron):
    """The large-batch kernel variant (4 anchor sets per CTA, no GT split) against the oracle on a
    sample of the 256 images, bit-exact."""
    net, anchors = ron
    boxes, labels, counts = synth.make_gt_batch(2, 256, 1, 50)
    r = net.bboxes_encode_batch(labels, boxes, counts, anchors, 0.56, 0.3, want_matched=True)
    _, (enc, cor, inside) = _oracle_tables(O.RON320)
    for b in list(range(0, 256, 13)) + [255]:
        o = O.encode_image(labels[b, :counts[b]], boxes[b, :counts[b]], enc, cor, inside, 0.56, 0.3)
        eq(r['matched'][b], o['matched'].astype(np.int32), 'matched[%d]' % b)
        eq(r['labels'][b], o['labels'], 'labels[%d]' % b)
        eq(r['scores'][b], o['scores'], 'scores[%d]' % b)
        eq(r['loc'][b], o['loc'], 'loc[%d]' % b)


@pytest.mark.parametrize('batch', [3, 160])
def test_encode_extreme_coordinates(ron, batch):
    """GT boxes with tiny coordinates / tiny sides / zero area (outside the range where the inline
    division sequence is proven exact: the kernel must switch to IEEE division) and duplicated GT
    boxes (exact ties between GT rows), both kernel variants."""
    net, anchors = ron
    boxes, labels, counts = synth.make_gt_batch(7, batch, 3, 12)
    boxes = boxes.copy()
    for b in range(batch):
        k = b % 5
        if k == 4:
            boxes[b, 0] = [-0.3, -0.2, 1.4, 1.3]                 # sides > 1: the saturating clamp must not be used
        elif k == 0:
            boxes[b, 0] = [1e-7, 3e-6, 0.4, 0.5]                 # tiny non-zero corner
        elif k == 1:
            boxes[b, 1] = [0.25, 0.25, 0.25 + 1e-6, 0.75]        # sliver
        elif k == 2:
            boxes[b, 2] = boxes[b, 0]                            # duplicate GT: lower index must win
        else:
            boxes[b, 1] = [0.5, 0.5, 0.5, 0.5]                   # zero area
    r = net.bboxes_encode_batch(labels, boxes, counts, anchors, 0.5, 0.3, want_matched=True)
    _, (enc, cor, inside) = _oracle_tables(O.RON320)
    for b in range(0, batch, max(1, batch // 12)):
        with np.errstate(all='ignore'):
            o = O.encode_image(labels[b, :counts[b]], boxes[b, :counts[b]], enc, cor, inside, 0.5, 0.3)
        eq(r['matched'][b], o['matched'].astype(np.int32), 'matched[%d]' % b)
        eq(r['labels'][b], o['labels'], 'labels[%d]' % b)
        eq(r['scores'][b], o['scores'], 'scores[%d]' % b)
        eq(r['loc'][b], o['loc'], 'loc[%d]' % b)


def test_caller_supplied_anchors_are_used_as_given(ron):
    """The reference always computes with the anchors it is handed.  A plain list of (y, x, h, w) arrays, a copied list,
    or a list whose arrays were edited after RONNet.anchors returned it must not be replaced by the cached default set."""
    import torch
    net, anchors = ron
    spec = O.RON320
    moved = []
    for (y, x, h, w) in anchors:
        moved.append((y + np.float32(0.013), x - np.float32(0.007), (h * np.float32(1.1)).astype(np.float32), w))
    enc, cor, inside = O.encode_anchor_tables(moved, spec.img_shape, spec.allowed_borders)
    boxes, labels, counts = synth.make_gt_batch(2, 1, 9, 9)
    want = O.encode_image(labels[0, :9], boxes[0, :9], enc, cor, inside, 0.5, 0.3)
    edited = type(anchors)(moved)                       # an AnchorList that still carries the default handle ...
    edited.anchor_set, edited.fingerprint = anchors.anchor_set, anchors.fingerprint   # ... and its stale fingerprint
    for what, a in (('plain list', list(moved)), ('edited AnchorList', edited)):
        lab, loc, sco, box = net.bboxes_encode(labels[0, :9], boxes[0, :9], a, positive_threshold=0.5, ignore_threshold=0.3)
        eq(torch.cat([t.reshape(-1) for t in lab]), want['labels'], what + ': labels')
        eq(torch.cat([t.reshape(-1) for t in sco]), want['scores'], what + ': scores')
        eq(torch.cat([t.reshape(-1, 4) for t in loc]), want['loc'], what + ': loc')
        eq(torch.cat([t.reshape(-1, 4) for t in box]), cor, what + ': anchor corner boxes')
    r = net.bboxes_encode_batch(labels, boxes, counts, list(moved), 0.5, 0.3)
    eq(r['labels'][0], want['labels'], 'batched form with a plain list')
    # decode with the moved anchors
    loc_in = np.random.Generator(np.random.PCG64(3)).normal(0, 0.3, size=(1, 21250, 4)).astype(np.float32)
    ls = anchors.anchor_set.layers
    per_layer = [loc_in[:, o:o + H * W * A].reshape(1, H, W, A, 4) for (H, W, A, o) in ls]
    got = net.bboxes_decode(per_layer, list(moved))
    eq(torch.cat([t.reshape(1, -1, 4) for t in got], 1)[0], O.decode(loc_in[0], O.flat_decode_anchors(moved)), 'decode')
    # the untouched list still takes the fast path (its own handle)
    assert net._resolve(anchors) is anchors.anchor_set and net._resolve(list(anchors)) is None


def test_two_graph_captures_do_not_share_a_workspace(ron):
    """Two CUDA-graph captures of the same encode shape, replayed in the opposite order: each capture owns a workspace
    that its own graph zeroes (a cached buffer allocated inside the first capture would be zeroed by graph 1 only)."""
    import torch
    from ron_tensorflow_b200 import core
    net, anchors = ron
    aset = anchors.anchor_set
    ba, la, ca = synth.make_gt_batch(2, 4, 1, 30, g_max=30)
    bb, lb, cb = synth.make_gt_batch(2, 4, 1, 30, g_max=30, first_image=50)
    da = [torch.from_numpy(x).cuda() for x in (ba, la, ca)]
    db = [torch.from_numpy(x).cuda() for x in (bb, lb, cb)]
    fn = lambda b, l, c: core.match_encode(aset, b, l, c, 0.5, 0.3, want_matched=True)
    core._ws_cache.clear()
    g1 = core.Graphed(fn, *da)
    g2 = core.Graphed(fn, *db)
    r2 = {k: v.clone() for k, v in g2.replay().items()}
    r1 = {k: v.clone() for k, v in g1.replay().items()}
    r2b = g2.replay()
    w1, w2 = fn(*da), fn(*db)
    for k in ('labels', 'loc', 'scores', 'matched'):
        eq(r1[k], w1[k], 'graph 1 ' + k)
        eq(r2[k], w2[k], 'graph 2 (replayed first) ' + k)
        eq(r2b[k], w2[k], 'graph 2 again ' + k)
