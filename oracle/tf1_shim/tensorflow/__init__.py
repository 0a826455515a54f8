"""A NumPy emulation of the handful of TensorFlow-1.x ops that the hot path of
HiKapok/RON_Tensorflow calls.  TEST INFRASTRUCTURE ONLY (see oracle/ron_oracle.py).

Purpose: TensorFlow 1.x cannot be installed in the build container (Python 3.12, no
wheel, no network), yet every hot-path module of the reference imports ``tensorflow``
at module scope.  Putting this directory first on ``sys.path`` lets the reference's
*unmodified* Python (``nets/ssd_common.py``, ``tf_extended/bboxes.py``,
``nets/ron_vgg_320.py`` ...) execute eagerly, so ``tests/golden/make_golden.py`` can
record its outputs as golden vectors.  What is pinned that way: the reference's
operand order, composition, tie-breaking and thresholds.  What is NOT: the TF
runtime's own kernels.  Each op below states the TF semantics it restates:

* one float32 rounding per op, no fusion (NumPy ufuncs on float32 arrays);
* Python scalars / ndarrays are converted to the dtype of the first Tensor operand
  (``ops.convert_to_tensor(y, dtype=x.dtype)``), e.g. a float64 ndarray compared with
  a float32 tensor is cast to float32 first;
* ``argmax`` returns the first maximal index; ``nn.top_k`` lists lower indices first
  among equal values; ``one_hot`` is all-off for out-of-range indices;
  ``boolean_mask`` keeps order;
* ``exp``/``log`` are the correctly rounded float32 results (definitional).

It is not a general TensorFlow replacement and never ships with the product.
"""
import os
import sys
import types
import importlib.abc
import importlib.machinery

import builtins as _b

import numpy as np

__version__ = '1.5.0-numpy-shim'


# --------------------------------------------------------------------------- #
# dtypes
# --------------------------------------------------------------------------- #
class DType(object):
    def __init__(self, name, np_dtype):
        self.name = name
        self.as_numpy_dtype = np_dtype
        self.base_dtype = self
        if np.issubdtype(np_dtype, np.integer):
            self.max = int(np.iinfo(np_dtype).max)
            self.min = int(np.iinfo(np_dtype).min)
        elif np.issubdtype(np_dtype, np.floating):
            self.max = float(np.finfo(np_dtype).max)
            self.min = float(np.finfo(np_dtype).min)

    def __eq__(self, other):
        return isinstance(other, DType) and other.name == self.name

    def __hash__(self):
        return hash(self.name)

    def __repr__(self):
        return 'tf.' + self.name


float32 = DType('float32', np.float32)
float64 = DType('float64', np.float64)
int8 = DType('int8', np.int8)
int32 = DType('int32', np.int32)
int64 = DType('int64', np.int64)
uint8 = DType('uint8', np.uint8)
bool = DType('bool', np.bool_)
_ALL = [float32, float64, int8, int32, int64, uint8, bool]
_pybool = _b.bool


def _dt(np_dtype):
    for d in _ALL:
        if np.dtype(d.as_numpy_dtype) == np.dtype(np_dtype):
            return d
    raise TypeError('unsupported dtype %r' % (np_dtype,))


def _npdt(d):
    if d is None:
        return None
    if isinstance(d, DType):
        return d.as_numpy_dtype
    return np.dtype(d).type


# --------------------------------------------------------------------------- #
# Tensor
# --------------------------------------------------------------------------- #
class TensorShape(object):
    def __init__(self, dims):
        self._d = [int(x) for x in dims]

    def is_fully_defined(self):
        return True

    def as_list(self):
        return list(self._d)

    def with_rank(self, r):
        assert len(self._d) == r
        return self

    def is_compatible_with(self, other):
        return self._d == other._d

    def __len__(self):
        return len(self._d)

    def __getitem__(self, i):
        return self._d[i]


class Tensor(object):
    __array_ufunc__ = None          # make ndarray <op> Tensor defer to Tensor.__r<op>__
    __array_priority__ = 1000

    def __init__(self, a):
        self.a = np.asarray(a)

    @property
    def dtype(self):
        return _dt(self.a.dtype)

    @property
    def shape(self):
        return TensorShape(self.a.shape)

    def get_shape(self):
        return TensorShape(self.a.shape)

    def numpy(self):
        return self.a

    def __repr__(self):
        return 'Tensor(%r)' % (self.a,)

    def __getitem__(self, idx):
        def fix(i):
            if isinstance(i, Tensor):
                return i.a if i.a.ndim else int(i.a)
            return i
        if isinstance(idx, _b.tuple):
            idx = _b.tuple(fix(i) for i in idx)
        else:
            idx = fix(idx)
        return Tensor(self.a[idx])

    def __iter__(self):
        for i in _b.range(self.a.shape[0]):
            yield Tensor(self.a[i])

    def __len__(self):
        return self.a.shape[0]

    def __bool__(self):
        return _pybool(self.a)

    __nonzero__ = __bool__

    def __index__(self):
        return int(self.a)

    def __int__(self):
        return int(self.a)

    def __float__(self):
        return float(self.a)

    def __hash__(self):
        return id(self)

    # arithmetic (math_ops.binary_op_wrapper: y -> convert_to_tensor(y, dtype=x.dtype))
    def __add__(self, o): return _bin(np.add, self, o)
    def __radd__(self, o): return _bin(np.add, o, self)
    def __sub__(self, o): return _bin(np.subtract, self, o)
    def __rsub__(self, o): return _bin(np.subtract, o, self)
    def __mul__(self, o): return _bin(np.multiply, self, o)
    def __rmul__(self, o): return _bin(np.multiply, o, self)
    def __truediv__(self, o): return _bin(_truediv, self, o)
    def __rtruediv__(self, o): return _bin(_truediv, o, self)
    __div__ = __truediv__
    __rdiv__ = __rtruediv__
    def __neg__(self): return Tensor(np.negative(self.a))
    def __lt__(self, o): return _bin(np.less, self, o)
    def __le__(self, o): return _bin(np.less_equal, self, o)
    def __gt__(self, o): return _bin(np.greater, self, o)
    def __ge__(self, o): return _bin(np.greater_equal, self, o)
    def __and__(self, o): return _bin(np.logical_and, self, o)
    def __or__(self, o): return _bin(np.logical_or, self, o)
    def __invert__(self): return Tensor(np.logical_not(self.a))


def _truediv(x, y):
    if np.issubdtype(x.dtype, np.integer):
        x = x.astype(np.float64)
        y = y.astype(np.float64)
    with np.errstate(divide='ignore', invalid='ignore'):
        return np.true_divide(x, y)


def _default(x):
    """Python values -> TF default dtypes (float -> float32, int -> int32)."""
    if isinstance(x, (list, _b.tuple)) and any(isinstance(e, Tensor) for e in x):
        x = [_v(e) for e in x]
    a = np.asarray(x)
    if isinstance(x, (np.ndarray, np.generic)):
        return a
    if a.dtype == np.float64:
        return a.astype(np.float32)
    if a.dtype == np.int64:
        return a.astype(np.int32)
    return a


def _v(x, dtype=None):
    """convert_to_tensor -> ndarray.  ``dtype`` is a numpy dtype or None."""
    if isinstance(x, Tensor):
        if dtype is not None and np.dtype(dtype) != x.a.dtype:
            raise TypeError('Tensor dtype %s, expected %s' % (x.a.dtype, np.dtype(dtype)))
        return x.a
    if dtype is None:
        return _default(x)
    if isinstance(x, (list, _b.tuple)) and any(isinstance(e, Tensor) for e in x):
        x = [_v(e) for e in x]
    return np.asarray(x).astype(dtype)


def _pair(x, y):
    if isinstance(x, Tensor):
        return x.a, _v(y, x.a.dtype)
    if isinstance(y, Tensor):
        return _v(x, y.a.dtype), y.a
    xa = _default(x)
    return xa, _v(y, xa.dtype)


def _bin(fn, x, y):
    a, b = _pair(x, y)
    with np.errstate(all='ignore'):
        return Tensor(fn(a, b))


def convert_to_tensor(x, dtype=None, name=None):
    return Tensor(_v(x, _npdt(dtype)))


def constant(value, dtype=None, shape=None, name=None):
    a = _v(value, _npdt(dtype))
    if shape is not None:
        a = np.broadcast_to(a, shape).copy()
    return Tensor(a)


# --------------------------------------------------------------------------- #
# element-wise
# --------------------------------------------------------------------------- #
def maximum(x, y, name=None): return _bin(np.maximum, x, y)
def minimum(x, y, name=None): return _bin(np.minimum, x, y)
def add(x, y, name=None): return _bin(np.add, x, y)
def subtract(x, y, name=None): return _bin(np.subtract, x, y)
def multiply(x, y, name=None): return _bin(np.multiply, x, y)
def divide(x, y, name=None): return _bin(_truediv, x, y)
def truediv(x, y, name=None): return _bin(_truediv, x, y)
def equal(x, y, name=None): return _bin(np.equal, x, y)
def not_equal(x, y, name=None): return _bin(np.not_equal, x, y)
def greater(x, y, name=None): return _bin(np.greater, x, y)
def greater_equal(x, y, name=None): return _bin(np.greater_equal, x, y)
def less(x, y, name=None): return _bin(np.less, x, y)
def less_equal(x, y, name=None): return _bin(np.less_equal, x, y)
def logical_and(x, y, name=None): return _bin(np.logical_and, x, y)
def logical_or(x, y, name=None): return _bin(np.logical_or, x, y)
def logical_not(x, name=None): return Tensor(np.logical_not(_v(x)))
mul = multiply


def _f64_round(fn, x):
    a = _v(x)
    if a.dtype == np.float32:
        with np.errstate(all='ignore'):
            return Tensor(fn(a.astype(np.float64)).astype(np.float32))
    with np.errstate(all='ignore'):
        return Tensor(fn(a))


def log(x, name=None): return _f64_round(np.log, x)
def exp(x, name=None): return _f64_round(np.exp, x)


def sqrt(x, name=None):
    """IEEE square root is correctly rounded in every implementation: one float32 rounding."""
    return Tensor(np.sqrt(_v(x)))


def cast(x, dtype, name=None):
    return Tensor(_v(x).astype(_npdt(dtype)))


def to_float(x, name=None): return cast(x, float32)
def to_int64(x, name=None): return cast(x, int64)
def to_int32(x, name=None): return cast(x, int32)


def zeros_like(x, dtype=None, name=None):
    a = _v(x)
    return Tensor(np.zeros(a.shape, _npdt(dtype) or a.dtype))


def ones_like(x, dtype=None, name=None):
    a = _v(x)
    return Tensor(np.ones(a.shape, _npdt(dtype) or a.dtype))


def _shape_arg(s):
    if isinstance(s, Tensor):
        return _b.tuple(int(v) for v in np.atleast_1d(s.a))
    if isinstance(s, (list, _b.tuple)):
        return _b.tuple(int(v) for v in s)
    return (int(s),)


def zeros(shape, dtype=float32, name=None):
    return Tensor(np.zeros(_shape_arg(shape), _npdt(dtype)))


def ones(shape, dtype=float32, name=None):
    return Tensor(np.ones(_shape_arg(shape), _npdt(dtype)))


def clip_by_value(x, lo, hi, name=None):
    a = _v(x)
    return Tensor(np.minimum(np.maximum(a, _v(lo, a.dtype)), _v(hi, a.dtype)))


def where(condition, x=None, y=None, name=None):
    c = _v(condition).astype(np.bool_)
    if x is None and y is None:
        return Tensor(np.argwhere(c).astype(np.int64))
    a, b = _pair(x, y)
    if c.ndim == 1 and a.ndim > 1:
        c = c.reshape((-1,) + (1,) * (a.ndim - 1))
    return Tensor(np.where(c, a, b))


# --------------------------------------------------------------------------- #
# shapes
# --------------------------------------------------------------------------- #
def shape(x, name=None, out_type=int32):
    return Tensor(np.asarray(_v(x).shape, _npdt(out_type)))


def size(x, name=None, out_type=int32):
    return Tensor(np.asarray(_v(x).size, _npdt(out_type)))


def rank(x, name=None):
    return Tensor(np.asarray(_v(x).ndim, np.int32))


def reshape(x, shp, name=None):
    return Tensor(_v(x).reshape(_shape_arg(shp)))


def expand_dims(x, axis=None, name=None, dim=None):
    return Tensor(np.expand_dims(_v(x), axis if axis is not None else dim))


def squeeze(x, axis=None, name=None, squeeze_dims=None):
    ax = axis if axis is not None else squeeze_dims
    if isinstance(ax, list):
        ax = _b.tuple(ax)
    a = _v(x)
    if ax is not None:
        for d in (ax if isinstance(ax, _b.tuple) else (ax,)):
            if a.shape[d] != 1:
                raise ValueError('Can not squeeze dim[%d], expected a dimension of 1, got %d'
                                 % (d, a.shape[d]))
    return Tensor(np.squeeze(a, ax))


def transpose(x, perm=None, name=None):
    return Tensor(np.transpose(_v(x), perm))


def stack(values, axis=0, name=None):
    if isinstance(values, Tensor):
        return values
    ref = next((e for e in values if isinstance(e, Tensor)), None)
    if ref is None:
        return Tensor(np.stack([_default(e) for e in values], axis))
    return Tensor(np.stack([_v(e, ref.a.dtype) for e in values], axis))


def unstack(x, num=None, axis=0, name=None):
    a = _v(x)
    return [Tensor(np.take(a, i, axis)) for i in _b.range(a.shape[axis])]


def concat(values, axis, name=None):
    ref = next((e for e in values if isinstance(e, Tensor)), None)
    if ref is None:
        return Tensor(np.concatenate([_default(e) for e in values], axis))
    return Tensor(np.concatenate([_v(e, ref.a.dtype) for e in values], axis))


def split(value, num_or_size_splits, axis=0, num=None, name=None):
    a = _v(value)
    if isinstance(num_or_size_splits, Tensor):
        num_or_size_splits = [int(v) for v in num_or_size_splits.a]
    if isinstance(num_or_size_splits, (list, _b.tuple, np.ndarray)):
        sizes = [int(v) for v in num_or_size_splits]
        assert sum(sizes) == a.shape[axis]
        cuts = np.cumsum(sizes)[:-1]
        return [Tensor(p) for p in np.split(a, cuts, axis)]
    return [Tensor(p) for p in np.split(a, int(num_or_size_splits), axis)]


def pad(x, paddings, mode='CONSTANT', name=None, constant_values=0):
    assert mode == 'CONSTANT'
    p = _v(paddings)
    return Tensor(np.pad(_v(x), [(int(r[0]), int(r[1])) for r in p], mode='constant',
                         constant_values=constant_values))


def reverse(x, axis, name=None):
    return Tensor(np.flip(_v(x), _b.tuple(axis)))


def range(start, limit=None, delta=1, dtype=None, name=None):
    s = int(_v(start))
    if limit is None:
        s, l = 0, s
        dt = _v(start).dtype
    else:
        l = int(_v(limit))
        dt = _v(start).dtype
    return Tensor(np.arange(s, l, int(delta)).astype(_npdt(dtype) or dt))


# --------------------------------------------------------------------------- #
# reductions / indexing
# --------------------------------------------------------------------------- #
def _axis(axis):
    if isinstance(axis, Tensor):
        return int(axis.a)
    if isinstance(axis, list):
        return _b.tuple(axis)
    return axis


def reduce_max(x, axis=None, keep_dims=False, name=None, keepdims=None):
    return Tensor(np.max(_v(x), axis=_axis(axis), keepdims=_pybool(keep_dims or keepdims)))


def reduce_min(x, axis=None, keep_dims=False, name=None, keepdims=None):
    return Tensor(np.min(_v(x), axis=_axis(axis), keepdims=_pybool(keep_dims or keepdims)))


def reduce_sum(x, axis=None, keep_dims=False, name=None, keepdims=None):
    a = _v(x)
    return Tensor(np.sum(a, axis=_axis(axis), keepdims=_pybool(keep_dims or keepdims)).astype(a.dtype))


def count_nonzero(x, axis=None, keep_dims=False, dtype=int64, name=None):
    return Tensor(np.asarray(np.count_nonzero(_v(x), axis=_axis(axis))).astype(_npdt(dtype)))


def argmax(x, axis=None, name=None, dimension=None, output_type=int64):
    ax = axis if axis is not None else dimension
    return Tensor(np.argmax(_v(x), axis=_axis(ax) if ax is not None else 0).astype(_npdt(output_type)))


def cumsum(x, axis=0, exclusive=False, reverse=False, name=None):
    assert not exclusive and not reverse
    a = _v(x)
    return Tensor(np.cumsum(a, axis=axis).astype(a.dtype))


def add_n(inputs, name=None):
    acc = inputs[0]
    for t in inputs[1:]:
        acc = acc + t
    return acc if isinstance(acc, Tensor) else Tensor(_v(acc))


def gather(params, indices, validate_indices=None, name=None, axis=0):
    return Tensor(np.take(_v(params), _v(indices), axis=axis))


def gather_nd(params, indices, name=None):
    p = _v(params)
    i = _v(indices)
    return Tensor(p[_b.tuple(i[..., k] for k in np.arange(i.shape[-1]))])


def boolean_mask(x, mask, name=None):
    return Tensor(_v(x)[_v(mask).astype(np.bool_)])


def one_hot(indices, depth, on_value=None, off_value=None, axis=None, dtype=None, name=None):
    idx = _v(indices)
    depth = int(_v(depth))
    if dtype is None:
        dtype = convert_to_tensor(on_value).dtype if on_value is not None else float32
    npd = _npdt(dtype)
    on = np.asarray(1 if on_value is None else _v(on_value)).astype(npd)
    off = np.asarray(0 if off_value is None else _v(off_value)).astype(npd)
    hit = (idx[..., None] == np.arange(depth))          # out-of-range indices: all off
    out = np.where(hit, on, off).astype(npd)
    if axis is not None and axis != -1:
        out = np.moveaxis(out, -1, axis)
    return Tensor(out)


def tuple(tensors, name=None, control_inputs=None):
    return list(tensors)


# --------------------------------------------------------------------------- #
# control flow
# --------------------------------------------------------------------------- #
def while_loop(cond, body, loop_vars, shape_invariants=None, parallel_iterations=10,
               back_prop=True, swap_memory=False, name=None):
    v = list(loop_vars)
    while _pybool(_v(cond(*v))):
        v = list(body(*v))
    return v


def cond(pred, true_fn=None, false_fn=None, name=None, fn1=None, fn2=None):
    t = true_fn or fn1
    f = false_fn or fn2
    return t() if _pybool(_v(pred)) else f()


def map_fn(fn, elems, dtype=None, parallel_iterations=10, back_prop=True,
           swap_memory=False, infer_shape=True, name=None):
    multi = isinstance(elems, (list, _b.tuple))
    first = elems[0] if multi else elems
    n = _v(first).shape[0]
    outs = []
    for i in _b.range(n):
        arg = type(elems)(Tensor(_v(e)[i]) for e in elems) if multi else Tensor(_v(elems)[i])
        outs.append(fn(arg))
    if isinstance(outs[0], (list, _b.tuple)):
        k = len(outs[0])
        res = [Tensor(np.stack([_v(o[j]) for o in outs])) for j in _b.range(k)]
        return type(outs[0])(res) if isinstance(outs[0], _b.tuple) else res
    return Tensor(np.stack([_v(o) for o in outs]))


def scan(fn, elems, initializer=None, parallel_iterations=10, back_prop=True,
         swap_memory=False, infer_shape=True, name=None):
    a = _v(elems)
    acc = Tensor(a[0]) if initializer is None else initializer
    out = [acc] if initializer is None else []
    for i in _b.range(1 if initializer is None else 0, a.shape[0]):
        acc = fn(acc, Tensor(a[i]))
        out.append(acc)
    return Tensor(np.stack([_v(o) for o in out]))


class TensorArray(object):
    def __init__(self, dtype, size=None, dynamic_size=None, clear_after_read=None,
                 tensor_array_name=None, handle=None, flow=None, infer_shape=True,
                 element_shape=None, colocate_with_first_write_call=True, name=None):
        self._dtype = _npdt(dtype)
        self._items = [None] * int(_v(size))

    def write(self, index, value, name=None):
        self._items[int(_v(index))] = _v(value).astype(self._dtype)
        return self

    def stack(self, name=None):
        return Tensor(np.stack(self._items)) if self._items else Tensor(np.zeros((0,), self._dtype))


class _Scope(object):
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return 'scope'

    def __exit__(self, *a):
        return False


name_scope = _Scope
variable_scope = _Scope
device = _Scope
control_dependencies = _Scope


# hooks for the golden generator: ron_losses (nets/ron_vgg_320.py:635-771) returns nothing, so the masks it
# wraps in tf.stop_gradient and the losses it hands to tf.losses.add_loss are recorded here; the uniforms
# of tf.random_uniform come from ``_rng`` (set by the caller) and are recorded too.
_stop_gradient_log = []
_loss_log = []
_random_log = []
_rng = None


def stop_gradient(x, name=None):
    _stop_gradient_log.append(x)
    return x


def random_uniform(shape, minval=0, maxval=None, dtype=None, seed=None, name=None):
    if _rng is None:
        raise RuntimeError('tf-shim: set tensorflow._rng = numpy.random.Generator(...) before tf.random_uniform')
    shp = _shape_arg(shape)
    u = _rng.random(size=shp, dtype=np.float32)
    hi = np.float32(1.0 if maxval is None else maxval)
    lo = np.float32(minval)
    u = (u * (hi - lo) + lo).astype(np.float32)
    _random_log.append(u)
    return Tensor(u)


def abs(x, name=None):
    return Tensor(np.abs(_v(x)))


def reduce_mean(x, axis=None, keep_dims=False, name=None, keepdims=None):
    a = _v(x)
    # float64 accumulate, one rounding: the reduction ORDER of the real kernel is unspecified,
    # so means are compared with a tolerance, never bit for bit
    r = np.mean(a.astype(np.float64), axis=_axis(axis), keepdims=_pybool(keep_dims or keepdims))
    return Tensor(np.asarray(r).astype(a.dtype))


def _sparse_softmax_xent(_sentinel=None, labels=None, logits=None, name=None):
    z = _v(logits).astype(np.float64)
    l = _v(labels).astype(np.int64)
    m = z.max(-1, keepdims=True)
    lse = np.log(np.exp(z - m).sum(-1)) + m[..., 0]
    return Tensor((lse - np.take_along_axis(z, l[..., None], -1)[..., 0]).astype(np.float32))


def _add_loss(loss, loss_collection=None):
    _loss_log.append(loss)


def identity(x, name=None):
    return x


def Print(x, data, message=None, first_n=None, summarize=None, name=None):
    return x


# --------------------------------------------------------------------------- #
# tf.nn
# --------------------------------------------------------------------------- #
def _top_k(x, k=1, sorted=True, name=None):
    a = _v(x)
    k = int(_v(k))
    if k > a.shape[-1]:
        raise ValueError('input must have at least k columns')
    # descending by value, lower index first among equals (TopK op contract)
    order = np.argsort(-a.astype(np.float64), axis=-1, kind='stable')[..., :k]
    return Tensor(np.take_along_axis(a, order, -1)), Tensor(order.astype(np.int32))




# --------------------------------------------------------------------------- #
# tf.train.Example / tf.python_io.TFRecordWriter / tf.gfile: enough for the reference's
# datasets/pascalvoc_to_tfrecords.py to run unmodified and write real TFRecord files
# (public formats: protobuf wire encoding of Example, TFRecord framing with masked CRC-32C)
# --------------------------------------------------------------------------- #
def _pb_varint(v):
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7f
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _pb_len(field, payload):
    return _pb_varint((field << 3) | 2) + _pb_varint(len(payload)) + payload


class _PbList(object):
    def __init__(self, value=()):
        self.value = list(value)


class BytesList(_PbList):
    def SerializeToString(self):
        return b''.join(_pb_len(1, bytes(v)) for v in self.value)


class FloatList(_PbList):
    def SerializeToString(self):          # repeated float value = 1 [packed = true]
        return _pb_len(1, np.asarray(self.value, '<f4').tobytes()) if self.value else b''


class Int64List(_PbList):
    def SerializeToString(self):          # repeated int64 value = 1 [packed = true]
        return _pb_len(1, b''.join(_pb_varint(int(v)) for v in self.value)) if self.value else b''


class Feature(object):
    def __init__(self, bytes_list=None, float_list=None, int64_list=None):
        self.bytes_list, self.float_list, self.int64_list = bytes_list, float_list, int64_list

    def SerializeToString(self):
        for field, v in ((1, self.bytes_list), (2, self.float_list), (3, self.int64_list)):
            if v is not None:
                return _pb_len(field, v.SerializeToString())
        return b''


class Features(object):
    def __init__(self, feature=None):
        self.feature = dict(feature or {})

    def SerializeToString(self):          # map<string, Feature> feature = 1; deterministic: sorted by key
        out = b''
        for k in sorted(self.feature):
            out += _pb_len(1, _pb_len(1, k.encode('utf-8')) + _pb_len(2, self.feature[k].SerializeToString()))
        return out


class Example(object):
    def __init__(self, features=None):
        self.features = features

    def SerializeToString(self):
        return _pb_len(1, self.features.SerializeToString())


def _crc32c(data):
    c = 0xffffffff
    for b in data:
        c ^= b
        for _ in _b.range(8):
            c = (c >> 1) ^ (0x82f63b78 if c & 1 else 0)
    return c ^ 0xffffffff


def _masked_crc(data):
    c = _crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xa282ead8) & 0xffffffff


class _TFRecordWriter(object):
    def __init__(self, path, options=None):
        self._f = open(path, 'wb')

    def write(self, record):
        import struct
        head = struct.pack('<Q', len(record))
        self._f.write(head + struct.pack('<I', _masked_crc(head)) + record + struct.pack('<I', _masked_crc(record)))

    def close(self):
        self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
        return False


class _FastGFile(object):
    def __init__(self, name, mode='r'):
        self._f = open(name, mode)

    def read(self, n=-1):
        return self._f.read(n)

    def close(self):
        self._f.close()


# --------------------------------------------------------------------------- #
# permissive stubs for everything else (slim, contrib, flags, ...)
# --------------------------------------------------------------------------- #
class _Stub(object):
    def __init__(self, name='stub'):
        object.__setattr__(self, '_n', name)

    def __getattr__(self, k):
        if k.startswith('__') and k.endswith('__'):
            raise AttributeError(k)
        return _Stub(self._n + '.' + k)

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k and not isinstance(a[0], _Stub):
            return a[0]                              # used as a decorator
        return _Stub(self._n + '()')

    def __iter__(self):
        return iter(())

    def __repr__(self):
        return '<tf-shim stub %s>' % self._n


class _StubModule(types.ModuleType):
    """tensorflow.a.b.c: attributes resolve to this package's ops first (so
    ``math_ops.greater`` is ``tf.greater``), otherwise to nested stub modules that are
    also callable (decorators such as ``add_arg_scope`` pass their function through)."""
    def __getattr__(self, k):
        if k.startswith('__') and k.endswith('__'):
            raise AttributeError(k)
        me = sys.modules[__name__]
        if k in me.__dict__:
            return me.__dict__[k]
        full = self.__name__ + '.' + k
        m = sys.modules.get(full)
        if m is None:
            m = _StubModule(full)
            m.__path__ = []
            sys.modules[full] = m
        return m

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k and not isinstance(a[0], (_Stub, _StubModule)):
            return a[0]
        return _Stub(self.__name__ + '()')

    def __iter__(self):
        return iter(())


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.startswith('tensorflow.'):
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


if not any(isinstance(f, _Finder) for f in sys.meta_path):
    sys.meta_path.insert(0, _Finder())

nn = _StubModule('tensorflow.nn')
nn.top_k = _top_k
nn.sparse_softmax_cross_entropy_with_logits = _sparse_softmax_xent
losses = _StubModule('tensorflow.losses')
losses.add_loss = _add_loss
contrib = _Stub('tensorflow.contrib')
app = _Stub('tensorflow.app')
logging = _Stub('tensorflow.logging')
summary = _Stub('tensorflow.summary')
train = _StubModule('tensorflow.train')
train.Example, train.Features, train.Feature = Example, Features, Feature
train.BytesList, train.FloatList, train.Int64List = BytesList, FloatList, Int64List
python_io = _StubModule('tensorflow.python_io')
python_io.TFRecordWriter = _TFRecordWriter
gfile = _StubModule('tensorflow.gfile')
gfile.FastGFile = _FastGFile
gfile.Exists = os.path.exists
gfile.MakeDirs = lambda p: os.makedirs(p, exist_ok=True)
image = _Stub('tensorflow.image')
layers = _Stub('tensorflow.layers')
GraphKeys = _Stub('tensorflow.GraphKeys')
