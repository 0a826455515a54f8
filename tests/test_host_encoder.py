"""Host-buffer path of match + encode (core.HostEncoder): sparse device->host packets applied to pinned host arrays
must reproduce the dense device results exactly, across reused slots, through a packet overflow and back."""
import ctypes

import numpy as np
import pytest

from ron_tensorflow_b200 import synth
from _util import need_cuda, eq


def _packet(ffi, cap, rows):
    import torch
    total = int(ffi.lib().ronk_sparse_rows_packet_bytes(cap))
    voff = total - 16 * cap
    buf = np.zeros(total, np.uint8)
    buf[:4].view(np.int32)[0] = len(rows)
    n = min(len(rows), cap)
    buf[16:16 + 4 * n].view(np.int32)[:] = [i for i, _ in rows[:n]]
    buf[voff:voff + 16 * n].view(np.float32)[:] = np.array([v for _, v in rows[:n]], np.float32).reshape(-1)
    return torch.from_numpy(buf)


def test_host_rows_apply_is_plain_host_code():
    """No GPU involved: write, clear-and-rewrite through the previous packet, overflow is refused untouched."""
    import torch
    from ron_tensorflow_b200 import core, _ffi
    loc = torch.zeros((100, 4), dtype=torch.float32)
    p1 = _packet(_ffi, 4, [(3, [1, 2, 3, 4]), (99, [-1, 0, 0.5, 9])])
    assert core.host_rows_apply(p1, None, 4, loc)
    want = np.zeros((100, 4), np.float32); want[3] = [1, 2, 3, 4]; want[99] = [-1, 0, 0.5, 9]
    assert np.array_equal(loc.numpy(), want)
    p2 = _packet(_ffi, 4, [(4, [5, 5, 5, 5])])
    assert core.host_rows_apply(p2, p1, 4, loc)
    want[:] = 0; want[4] = 5
    assert np.array_equal(loc.numpy(), want)
    big = _packet(_ffi, 4, [(i, [1, 1, 1, 1]) for i in range(5)])
    assert not core.host_rows_apply(big, p2, 4, loc)                         # 5 rows > capacity 4
    assert np.array_equal(loc.numpy(), want)


@pytest.mark.gpu
def test_host_encoder_matches_dense_results_across_slots_and_overflow():
    need_cuda()
    import torch
    from ron_tensorflow_b200 import core
    from ron_tensorflow_b200.nets import ron_vgg_320
    net = ron_vgg_320.RONNet()
    anchors = net.anchors(net.params.img_shape)
    B, G = 8, 50
    batches = [synth.make_gt_batch(2, B, 1, 50, g_max=G, first_image=k * B) for k in range(5)]
    light = synth.make_gt_batch(2, B, 1, 1, g_max=G, first_image=900)       # one GT per image: a short packet

    def dense(b):
        r = net.bboxes_encode_batch(b[1], b[0], b[2], anchors, 0.56, 0.3)
        return {k: v.cpu().numpy() for k, v in r.items()}

    def check(h, b, what):
        want = dense(b)
        for k in ('labels', 'loc', 'scores'):
            eq(h[k].numpy(), want[k], '%s %s' % (what, k))

    aset = anchors.anchor_set
    enc = net.host_encoder(B, G, anchors, slots=2, positive_threshold=0.56, ignore_threshold=0.3)
    pin = lambda b: [torch.from_numpy(x).pin_memory() for x in b]
    enc.submit(0, *pin(batches[0]))
    for k in range(1, 5):                                                    # pipelined: slots alternate, each reused
        enc.submit(k % 2, *pin(batches[k]))
        check(enc.collect((k - 1) % 2), batches[k - 1], 'batch %d' % (k - 1))
    check(enc.collect(0), batches[4], 'batch 4')
    assert enc.d2h_bytes_per_step < 0.5 * B * aset.N * 28
    # capacities that the 1-GT batch fits and the 1..50-GT batches overflow: sparse -> dense fallback -> sparse again
    nz = lambda b: int((dense(b)['loc'] != 0).any(-1).sum())
    n_light, n_heavy = nz(light), min(nz(batches[0]), nz(batches[1]))
    assert 32 < n_light < n_heavy
    frac = (n_light + n_heavy) / 2.0 / (B * aset.N)
    small = core.HostEncoder(aset, B, G, 1, 0.56, 0.3, net.params.prior_scaling, loc_fraction=frac)
    for name, b in (('light', light), ('heavy (overflow)', batches[0]), ('light again', light), ('heavy again', batches[1]),
                    ('light after two dense', light)):
        small.submit(0, *pin(b))
        check(small.collect(0), b, name)
    with pytest.raises(RuntimeError):
        small.collect(0)


@pytest.mark.gpu
def test_host_encoder_sparse_labels():
    """The label packets (non-zero labels as (index, value) pairs, applied by host threads) against the dense results:
    a zero-area GT box (its all-zero overlap row force-matches anchor 0, an anchor outside the border mask), a label packet that overflows
    (dense fallback for that batch, sparse again afterwards), an image without GT boxes, labels beyond int32, and the
    dense labels of the first version (sparse_labels=False)."""
    need_cuda()
    import torch
    from ron_tensorflow_b200 import core
    from ron_tensorflow_b200.nets import ron_vgg_320
    net = ron_vgg_320.RONNet()
    anchors = net.anchors(net.params.img_shape)
    aset = anchors.anchor_set
    B, G = 6, 20
    normal = synth.make_gt_batch(2, B, 1, 20, g_max=G, first_image=40)
    corner = tuple(x.copy() for x in synth.make_gt_batch(2, B, 2, 20, g_max=G, first_image=80))
    corner[0][:, 1] = [0.5, 0.5, 0.5, 0.5]                   # zero area: overlaps nothing, anchor 0 is forced with score 0
    wide = tuple(x.copy() for x in synth.make_gt_batch(2, B, 1, 20, g_max=G, first_image=120))
    wide[1][:, 0] = 1 << 40                                   # does not fit int32
    empty = tuple(x.copy() for x in normal)
    empty[2][2] = 0                                           # an image without ground truth: everything background
    light = synth.make_gt_batch(2, B, 1, 1, g_max=G, first_image=900)

    def dense(b):
        r = net.bboxes_encode_batch(b[1], b[0], b[2], anchors, 0.5, 0.3)
        return {k: v.cpu().numpy() for k, v in r.items()}

    pin = lambda b: [torch.from_numpy(x).pin_memory() for x in b]
    assert dense(corner)['labels'][:, 0].tolist() == corner[1][:, 1].tolist(), 'anchor 0 must carry the forced label'
    nzl = lambda b: int((dense(b)['labels'] != 0).sum())
    assert nzl(light) < nzl(normal)
    tight = (nzl(light) + nzl(normal)) / 2.0 / (B * aset.N)   # fits the 1-GT batch, overflows the others
    for sparse, frac in ((True, 0.25), (True, tight), (False, 0.25)):
        enc = core.HostEncoder(aset, B, G, 2, 0.5, 0.3, net.params.prior_scaling, sparse_labels=sparse, label_fraction=frac,
                               host_threads=3)
        if sparse and frac == 0.25:
            assert enc.d2h_bytes_per_step < 0.35 * B * aset.N * 28
        for k, (name, b) in enumerate((('normal', normal), ('corner', corner), ('light', light), ('wide labels', wide),
                                       ('normal again', normal), ('empty image', empty), ('light again', light),
                                       ('corner again', corner))):
            enc.submit(k % 2, *pin(b))
            h = enc.collect(k % 2)
            want = dense(b)
            for key in ('labels', 'loc', 'scores'):
                eq(h[key].numpy(), want[key], '%s (sparse=%s, %.3f) %s' % (name, sparse, frac, key))


def test_host_targets_apply_is_plain_host_code():
    """ronk_host_targets_apply needs no GPU: previous entries zeroed, new ones written, overflow reported per tensor."""
    import ctypes
    from ron_tensorflow_b200 import _ffi as ffi
    lib = ffi.lib()
    T, lcap, rcap = 200, 8, 4

    def lab_packet(pairs, flag=0):
        buf = np.zeros(int(lib.ronk_sparse_labels_packet_bytes(lcap)), np.uint8)
        buf[:8].view(np.int32)[:] = [len(pairs), flag]
        n = min(len(pairs), lcap)
        buf[16:16 + 4 * n].view(np.int32)[:] = [i for i, _ in pairs[:n]]
        buf[16 + 4 * lcap:16 + 4 * lcap + 4 * n].view(np.int32)[:] = [v for _, v in pairs[:n]]
        return buf

    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    labels = np.zeros(T, np.int64)
    loc = np.zeros((T, 4), np.float32)
    r1 = _packet(ffi, rcap, [(3, [1, 2, 3, 4]), (150, [5, 6, 7, 8])]).numpy()
    l1 = lab_packet([(0, 7), (3, -1), (199, 20)])
    for threads in (1, 4):
        assert lib.ronk_host_targets_apply(vp(r1), None, rcap, vp(loc), vp(l1), None, lcap, vp(labels), threads) == 0
        assert labels[[0, 3, 199]].tolist() == [7, -1, 20] and int((labels != 0).sum()) == 3
        assert loc[150].tolist() == [5, 6, 7, 8] and int((loc != 0).any(1).sum()) == 2
    r2 = _packet(ffi, rcap, [(9, [9, 9, 9, 9])]).numpy()
    l2 = lab_packet([(3, 2), (50, -1)])
    assert lib.ronk_host_targets_apply(vp(r2), vp(r1), rcap, vp(loc), vp(l2), vp(l1), lcap, vp(labels), 2) == 0
    assert np.flatnonzero(labels).tolist() == [3, 50] and labels[3] == 2 and np.flatnonzero(loc.any(1)).tolist() == [9]
    big = lab_packet([(i, 1) for i in range(lcap + 1)])
    before = labels.copy()
    assert lib.ronk_host_targets_apply(vp(r1), vp(r2), rcap, vp(loc), vp(big), vp(l2), lcap, vp(labels), 2) == 2   # labels untouched
    assert np.array_equal(labels, before) and np.flatnonzero(loc.any(1)).tolist() == [3, 150]
    assert lib.ronk_host_targets_apply(vp(r1), None, rcap, vp(loc), vp(lab_packet([(1, 1)], flag=1)), None, lcap, vp(labels), 1) == 2
