import numpy as np
import pytest


def need_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    return torch


def eq(a, b, what=''):
    """bit-exact comparison (NaN == NaN, +0 == -0 numerically)."""
    import torch
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    if isinstance(b, torch.Tensor):
        b = b.detach().cpu().numpy()
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, '%s shape %s vs %s' % (what, a.shape, b.shape)
    if a.dtype.kind == 'f':
        ok = (a == b) | ((a != a) & (b != b))
    else:
        ok = a == b
    if not ok.all():
        bad = np.argwhere(~ok)
        raise AssertionError('%s: %d/%d mismatches, first at %s: got %r expected %r' % (
            what, bad.shape[0], ok.size, tuple(bad[0]), a[tuple(bad[0])], b[tuple(bad[0])]))


def close(a, b, rtol=1e-5, what=''):
    import torch
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    if isinstance(b, torch.Tensor):
        b = b.detach().cpu().numpy()
    assert a.shape == b.shape, '%s shape %s vs %s' % (what, a.shape, b.shape)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=1e-7, err_msg=what)
