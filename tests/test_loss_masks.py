"""RON loss example masks + smooth-L1 (reference nets/ron_vgg_320.py:686-764,
nets/custom_layers.py:31-50).  The golden fixture holds what the reference's own ``ron_losses`` wrapped in
tf.stop_gradient / handed to tf.losses.add_loss when executed over the TF-1 shim with recorded uniforms
(tests/golden/make_golden.py: gen_loss_masks)."""
import hashlib

import numpy as np
import pytest

from oracle import ron_oracle as O
from ron_tensorflow_b200 import synth
from _util import need_cuda, eq

LS = [250, 1000, 4000, 16000]
MASKS = ('final_neg_mask_objness', 'objness_pred_label', 'cls_positive_mask', 'final_cls_neg_mask_objness')


def _flat_layers(x):
    """[B,N,...] -> the flat layer-major order of ron_losses (:660-675)."""
    return np.concatenate([t.reshape((-1,) + t.shape[2:]) for t in synth.split_layers(x, LS)])


def _inputs(g, tag):
    seed, batch = int(g[tag + '_cfg'][0]), int(g[tag + '_cfg'][1])
    logits, loc, obj_logits, obj_pred = synth.make_loss_inputs(seed, batch)
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    n = batch * sum(LS)
    r1 = rng.random(size=(n,), dtype=np.float32)
    r2 = rng.random(size=(n,), dtype=np.float32)
    if hashlib.sha256(r1.tobytes() + r2.tobytes()).digest() != g[tag + '_rand_sha'].tobytes():
        pytest.skip('numpy Generator stream differs from the one that made the fixture')
    return g[tag + '_gclasses'].astype(np.int64), _flat_layers(obj_pred), _flat_layers(loc), r1, r2


def _unpack(g, tag, name, n):
    return np.unpackbits(g[tag + '_' + name])[:n].astype(bool)


@pytest.mark.parametrize('tag', ['a', 'b'])
def test_oracle_matches_reference_ron_losses(golden, tag):
    g = golden('loss_masks')
    gcls, obj, loc, r1, r2 = _inputs(g, tag)
    m = O.ron_loss_masks(gcls, obj, r1, r2, 0.03, 3.)
    for k in MASKS:
        assert np.array_equal(m[k].astype(bool), _unpack(g, tag, k, gcls.shape[0])), k
    assert m['counts'][0] == _unpack(g, tag, 'objness_pred_label', gcls.shape[0]).sum()
    a, b = g[tag + '_sl1_pred'], g[tag + '_sl1_target']
    assert np.array_equal(O.modified_smooth_l1(a, b, sigma=3.), g[tag + '_sl1'])
    assert np.array_equal(O.modified_smooth_l1(a, b, 0.5, 2., sigma=1.), g[tag + '_sl1_w'])


def test_oracle_smooth_l1_and_degenerate_counts():
    x = np.array([-2., -1. / 9, -0.1, 0., 0.1, 1. / 9, 0.5, 3.], np.float32)
    out = O.modified_smooth_l1(x, np.zeros_like(x), sigma=3.)
    # |x| < 1/9: 4.5 x^2, else |x| - 1/18
    want = np.where(np.abs(x) < np.float32(1. / 9), (x * x) * np.float32(4.5), np.abs(x) - np.float32(0.5 / 9.))
    assert np.array_equal(out, want.astype(np.float32))
    # no positives: nothing is selected, the objectness label is all zero
    m = O.ron_loss_masks(np.zeros(100, np.int64), np.full(100, 0.5, np.float32), np.zeros(100, np.float32),
                         np.zeros(100, np.float32))
    assert not m['final_neg_mask_objness'].any() and not m['final_cls_neg_mask_objness'].any()
    # no negatives: safe_divide gives probability 0
    m = O.ron_loss_masks(np.ones(10, np.int64), np.full(10, 0.5, np.float32), np.zeros(10, np.float32), np.zeros(10, np.float32))
    assert m['final_neg_mask_objness'].all() and m['counts'].tolist() == [10., 0., 10., 0.]


@pytest.mark.gpu
@pytest.mark.parametrize('tag', ['a', 'b'])
def test_cuda_matches_reference_ron_losses(golden, tag):
    need_cuda()
    from ron_tensorflow_b200.nets import ron_vgg_320, custom_layers
    g = golden('loss_masks')
    gcls, obj, loc, r1, r2 = _inputs(g, tag)
    n = gcls.shape[0]
    m = ron_vgg_320.ron_loss_masks(gcls, obj, r1, r2, objness_threshold=0.03, negative_ratio=3.)
    for k in MASKS:
        eq(m[k].to('cpu').numpy().astype(bool), _unpack(g, tag, k, n), k)
    ref = O.ron_loss_masks(gcls, obj, r1, r2, 0.03, 3.)
    eq(m['counts'], ref['counts'], 'counts')
    # element-wise smooth-L1 against the reference's own function, bit for bit
    a, b = g[tag + '_sl1_pred'], g[tag + '_sl1_target']
    eq(custom_layers.modified_smooth_l1(a, b, sigma=3.), g[tag + '_sl1'], 'smooth l1')
    eq(custom_layers.modified_smooth_l1(a, b, 0.5, 2., 1.), g[tag + '_sl1_w'], 'smooth l1 weights')
    big = np.random.Generator(np.random.PCG64(5)).normal(0, 1, (100000, 4)).astype(np.float32)
    eq(custom_layers.modified_smooth_l1(big, big[::-1].copy(), sigma=3.), O.modified_smooth_l1(big, big[::-1], sigma=3.), 'smooth l1 big')
    # localisation loss: float reduction, tolerance 1e-5 relative (north_star)
    gl = np.random.Generator(np.random.PCG64(6)).normal(0, 0.5, loc.shape).astype(np.float32)
    got = float(ron_vgg_320.ron_localization_loss(loc, gl, m['cls_positive_mask']))
    want = float(O.ron_localization_loss(loc, gl, ref['cls_positive_mask']))
    assert abs(got - want) <= 1e-5 * abs(want), (got, want)
    # the fused form (masks + localisation term in one launch), twice: the workspace must come back zeroed
    for _ in range(2):
        f = ron_vgg_320.ron_loss_masks(gcls, obj, r1, r2, objness_threshold=0.03, localisations=loc, glocalisations=gl)
        assert abs(float(f['localization_loss']) - want) <= 1e-5 * abs(want)
        for k in MASKS:
            eq(f[k].to('cpu').numpy().astype(bool), _unpack(g, tag, k, n), k + ' (fused)')


@pytest.mark.gpu
def test_cuda_loss_masks_batch64_and_edge_cases():
    """Batch-64 sized input (1.36 M anchors) against the oracle, lists over layers, no positives / no negatives,
    device-drawn uniforms."""
    need_cuda()
    import torch
    from ron_tensorflow_b200.nets import ron_vgg_320
    rng = np.random.Generator(np.random.PCG64(77))
    n = 64 * 21250
    gcls = rng.choice(np.array([-1, 0, 0, 0, 0, 0, 0, 0, 3, 17], np.int64), size=n)
    obj = rng.random(n, dtype=np.float32) * np.float32(0.2)
    r1, r2 = rng.random(n, dtype=np.float32), rng.random(n, dtype=np.float32)
    m = ron_vgg_320.ron_loss_masks(gcls, obj, r1, r2, objness_threshold=0.1, negative_ratio=3.)
    ref = O.ron_loss_masks(gcls, obj, r1, r2, 0.1, 3.)
    for k in MASKS:
        eq(m[k].to('cpu').numpy().astype(bool), ref[k].astype(bool), k)
    eq(m['counts'], ref['counts'], 'counts')
    # views that are not 16-byte aligned take the scalar path; an odd length exercises the (n % 4) tail
    def off1(x):
        t = torch.from_numpy(np.concatenate([x[:1], x])).cuda()
        return t[1:]
    for m_ in (n - 1, 4099):
        mu = ron_vgg_320.ron_loss_masks(off1(gcls[:m_]), off1(obj[:m_]), off1(r1[:m_]), off1(r2[:m_]), objness_threshold=0.1)
        ma = ron_vgg_320.ron_loss_masks(gcls[:m_], obj[:m_], r1[:m_], r2[:m_], objness_threshold=0.1)
        ru = O.ron_loss_masks(gcls[:m_], obj[:m_], r1[:m_], r2[:m_], 0.1, 3.)
        for k in MASKS:
            eq(mu[k].to('cpu').numpy().astype(bool), ru[k].astype(bool), k + ' (unaligned)')
            eq(ma[k].to('cpu').numpy().astype(bool), ru[k].astype(bool), k + ' (odd length)')
    # lists over layers are flattened and concatenated
    cut = [1000, 50000, n - 51000]
    parts = lambda x: [torch.from_numpy(p) for p in np.split(x, np.cumsum(cut)[:-1])]
    m2 = ron_vgg_320.ron_loss_masks(parts(gcls), parts(obj), parts(r1), parts(r2), objness_threshold=0.1)
    assert torch.equal(m2['final_cls_neg_mask_objness'], m['final_cls_neg_mask_objness'])
    # no positives -> nothing selected and a zero localisation loss
    z = ron_vgg_320.ron_loss_masks(np.zeros(5000, np.int64), obj[:5000], r1[:5000], r2[:5000])
    assert not bool(z['final_neg_mask_objness'].any()) and z['counts'].tolist()[0] == 0.
    loc = rng.normal(0, 1, (5000, 4)).astype(np.float32)
    assert float(ron_vgg_320.ron_localization_loss(loc, loc * 0, z['cls_positive_mask'])) == 0.
    # device-drawn uniforms: positives always kept, selected negatives close to 3 x positives
    d = ron_vgg_320.ron_loss_masks(gcls, obj, objness_threshold=0.1, generator=torch.Generator(device='cuda').manual_seed(1))
    pos = gcls > 0
    sel = d['final_neg_mask_objness'].to('cpu').numpy()
    assert sel[pos].all() and not sel[gcls < 0].any()
    assert abs(sel[gcls == 0].sum() - 3 * pos.sum()) < 0.02 * 3 * pos.sum()


@pytest.mark.gpu
def test_localization_loss_and_smooth_l1_are_differentiable():
    """In the reference the localisation term is a training loss differentiated w.r.t. the network's localisations
    (glocalisations sits under tf.stop_gradient, nets/ron_vgg_320.py:760): loss.backward() must reach them, with the
    gradient of a plain float32 torch restatement of the same formula; modified_smooth_l1 likewise, element-wise."""
    need_cuda()
    import torch
    from ron_tensorflow_b200.nets import ron_vgg_320 as rv, custom_layers
    gen = torch.Generator(device='cuda').manual_seed(5)
    n = 5000
    loc = (torch.randn((n, 4), device='cuda', generator=gen) * 0.4).requires_grad_()
    gloc = torch.randn((n, 4), device='cuda', generator=gen) * 0.4
    mask = torch.rand((n,), device='cuda', generator=gen) < 0.2

    def torch_smooth(x, sigma):
        s2 = sigma * sigma
        return torch.where(x.abs() < 1. / s2, 0.5 * s2 * x * x, x.abs() - 0.5 / s2)

    loss = rv.ron_localization_loss(loc, gloc, mask)
    (3. * loss).backward()
    ref_in = loc.detach().clone().requires_grad_()
    ref = (1. / 3) * torch_smooth(ref_in - gloc, 3.)[mask].sum(-1).mean()
    (3. * ref).backward()
    assert abs(float(loss.detach()) - float(ref.detach())) <= 1e-5 * abs(float(ref.detach()))
    assert torch.allclose(loc.grad, ref_in.grad, rtol=1e-5, atol=1e-9)
    assert float(loc.grad[~mask].abs().sum()) == 0.
    # through ron_loss_masks (the fused launch yields the value; with grad-requiring localisations the differentiable form)
    labels = torch.randint(-1, 5, (n,), device='cuda', generator=gen)
    obj = torch.rand((n,), device='cuda', generator=gen)
    loc2 = loc.detach().clone().requires_grad_()
    out = rv.ron_loss_masks(labels, obj, torch.rand((n,), device='cuda', generator=gen), torch.rand((n,), device='cuda', generator=gen),
                            localisations=loc2, glocalisations=gloc)
    assert out['localization_loss'].requires_grad
    out['localization_loss'].backward()
    m2 = out['cls_positive_mask']
    ref2_in = loc.detach().clone().requires_grad_()
    if int(m2.sum()):
        ((1. / 3) * torch_smooth(ref2_in - gloc, 3.)[m2].sum(-1).mean()).backward()
        assert torch.allclose(loc2.grad, ref2_in.grad, rtol=1e-5, atol=1e-9)
    # element-wise smooth L1, gradient to both arguments
    a = (torch.randn((300, 4), device='cuda', generator=gen)).requires_grad_()
    b = (torch.randn((300, 4), device='cuda', generator=gen)).requires_grad_()
    w = torch.rand((300, 4), device='cuda', generator=gen)
    (custom_layers.modified_smooth_l1(a, b, 1., 1., sigma=2.) * w).sum().backward()
    a2, b2 = a.detach().clone().requires_grad_(), b.detach().clone().requires_grad_()
    (torch_smooth(a2 - b2, 2.) * w).sum().backward()
    assert torch.allclose(a.grad, a2.grad, rtol=1e-5, atol=1e-9) and torch.allclose(b.grad, b2.grad, rtol=1e-5, atol=1e-9)
