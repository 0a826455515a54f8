"""CPU-only: the C-ABI library loads and exports every symbol include/ronk.h declares, with the
argument counts the ctypes binding uses; without a GPU every compute call fails loudly."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, 'include', 'ronk.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    out = {}
    for m in re.finditer(r'\b(ronk_\w+)\s*\(([^;{]*?)\)\s*;', src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ('', 'void') else len([a for a in args.split(',')])
        out[m.group(1)] = n
    return out


@pytest.fixture(scope='module')
def ffi():
    from ron_tensorflow_b200 import _ffi, build
    build.build()               # no-op when up to date; nvcc cross-compiles without a GPU
    return _ffi


def test_header_and_binding_agree(ffi):
    decl = _header_functions()
    assert len(decl) >= 25
    assert set(decl) == set(ffi.SIGNATURES), set(decl) ^ set(ffi.SIGNATURES)
    for name, n in decl.items():
        assert len(ffi.SIGNATURES[name][1]) == n, name


def test_library_exports_every_symbol(ffi):
    lib = ffi.lib()
    for name in _header_functions():
        assert hasattr(lib, name), name
    assert lib.ronk_version() == 100
    assert lib.ronk_launch_count() >= 0


def test_argument_errors_without_gpu(ffi):
    import ctypes
    lib = ffi.lib()
    h = ctypes.c_void_p()
    rc = lib.ronk_anchors_create(7, 320, 320, 1, None, None, None, None, None, None, 0.5, None, ctypes.byref(h))
    assert rc == ffi.RONK_EINVAL and b'kind' in lib.ronk_last_error()
    with pytest.raises(ValueError):
        ffi.check(rc)
    assert lib.ronk_encode_workspace_bytes(64, 50) == 64 * 50 * 12 + 64 * 8     # 12 B per GT slot + per image a counter and an order slot
    assert lib.ronk_nms_workspace_bytes(10, 400) == 10 * 400 * 4
    rc = lib.ronk_nms_batch(None, None, 1, 1, 0.5, 1, 0, 1, None, None, None, None, None)
    assert rc == ffi.RONK_EINVAL


def test_no_cpu_fallback():
    """The product path must fail loudly without CUDA, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from ron_tensorflow_b200.nets import ron_vgg_320
    net = ron_vgg_320.RONNet()
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        net.anchors((320, 320))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'ron_tensorflow_b200')
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(d, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', txt, flags=re.M), os.path.join(d, f)
                assert 'tf1_shim' not in txt, os.path.join(d, f)
