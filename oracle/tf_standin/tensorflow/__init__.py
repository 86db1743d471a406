"""numpy stand-in for the subset of the TensorFlow API that CASAPose's post-network path uses.

TEST INFRASTRUCTURE (see oracle/__init__.py).  This is NOT TensorFlow and not part of the
product.  TensorFlow 2.9.1 cannot be installed in this image, so the reference's own source
files (/root/reference/casapose/pose_estimation/{ransac_voting,voting_layers_2d,
pose_evaluation,bpnp_layers}.py and casapose/utils/geometry_utils.py) are imported
UNMODIFIED with this module standing in for ``import tensorflow as tf``
(oracle/make_golden.py puts this directory in front of sys.path).  What that buys: the
golden vectors under tests/golden/ are produced by the reference's *code* — its op order,
axis conventions, flips, gates, tie rules and control flow — not by our restatement of it.
What it does not buy: TensorFlow's own kernels.  Each op below is the obvious numpy
equivalent with TensorFlow's dtype rules (python floats become float32, python ints
int32, integer reductions keep their dtype, one float32 rounding per op); powf, SVD and
the like are numpy's / LAPACK's, not Eigen's.  The order in which TensorFlow adds up a
float32 reduction or contraction is a property of its Eigen kernels (blocking, FMA,
thread count), so with ``ACCUMULATE = "float64"`` (default) ``reduce_sum`` / ``reduce_mean``
/ ``matmul`` on float32 operands accumulate in float64 and round once — the
order-independent value every float32 order approximates; ``ACCUMULATE = "native"`` keeps
numpy's own float32 order (pairwise sums, BLAS sgemm) and is used by the generator to
record how far a float32 order moves the result (4e-3 px on a 480x640 frame).

Two hooks exist for the generator script and nothing else:
  * ``random.uniform`` asks ``random.provider(shape, minval, maxval, dtype, map_index)``
    for its numbers — the script supplies the same explicit Philox streams the oracle and
    the CUDA path consume (tf.random.uniform's own stream is not reproducible anyway);
  * ``taps`` records the operands of ``argmax`` calls (the reference's vote-count tensor
    ``cur_inlier_counts`` only exists as that operand) together with the map_fn index.
"""
import builtins as _bi

import numpy as _np

__version__ = "0.0-numpy-standin"

float16 = _np.float16
float32 = _np.float32
float64 = _np.float64
int32 = _np.int32
int64 = _np.int64
uint8 = _np.uint8
bool = _np.bool_  # noqa: A001  (tf.bool)
newaxis = None

ACCUMULATE = "float64"

_map_index = []  # stack of tf.map_fn iteration indices (outermost first)
taps = []  # [(name, tuple(map_index), ndarray)]


def _t(x, dtype=None):
    """convert_to_tensor with TensorFlow's defaults: python float -> float32, python int -> int32."""
    if isinstance(x, _np.ndarray) or isinstance(x, _np.generic):
        return x if dtype is None else _np.asarray(x).astype(dtype, copy=False)
    if dtype is not None:
        return _np.asarray(x, dtype=dtype)
    a = _np.asarray(x)
    if a.dtype == _np.float64:
        return a.astype(_np.float32)
    if a.dtype == _np.int64:
        return a.astype(_np.int32)
    return a


def _like(x, other):
    """Python scalars / lists take the dtype of the tensor they are combined with."""
    if isinstance(x, (_np.ndarray, _np.generic)):
        return x
    if isinstance(other, (_np.ndarray, _np.generic)):
        return _np.asarray(x, dtype=_np.asarray(other).dtype)
    return _t(x)


def _pack(values):
    """Auto-packing of a python list that mixes tensors and python scalars."""
    ref = None
    for v in values:
        if isinstance(v, (_np.ndarray, _np.generic)):
            ref = v
            break
    return [_like(v, ref) if ref is not None else _t(v) for v in values]


def _keepint(x, fn, axis, keepdims=False):
    x = _np.asarray(x)
    if isinstance(axis, list):
        axis = tuple(axis)
    if x.dtype.kind in "iu" or x.dtype == _np.bool_:
        return fn(x, axis=axis, dtype=x.dtype, keepdims=keepdims)
    if x.dtype == _np.float32 and ACCUMULATE == "float64":
        return _np.asarray(fn(x, axis=axis, dtype=_np.float64, keepdims=keepdims)).astype(_np.float32)
    return fn(x, axis=axis, keepdims=keepdims)


# ---------------------------------------------------------------- graph plumbing
def function(fn=None, **_kw):
    if fn is not None:
        return fn
    return lambda f: f


def custom_gradient(f):
    def wrapper(*a, **k):
        out, _grad = f(*a, **k)
        return out

    return wrapper


def stop_gradient(x):
    return x


def numpy_function(func, inp, Tout):
    return _np.asarray(func(*[_np.asarray(i) for i in inp])).astype(Tout)


def map_fn(fn, elems, dtype=None, **_kw):
    n = len(elems[0]) if isinstance(elems, tuple) else len(elems)
    outs = []
    for i in _bi.range(n):
        _map_index.append(i)
        try:
            # TensorFlow tensors are immutable (``x *= y`` rebinds); hand out copies so that the reference's
            # in-place operators (ransac_voting.py:301, :306) cannot write through a view into the caller's data
            e = tuple(_np.array(x[i]) for x in elems) if isinstance(elems, tuple) else _np.array(elems[i])
            outs.append(_np.asarray(fn(e)))
        finally:
            _map_index.pop()
    out = _np.stack(outs, axis=0)
    return out.astype(dtype) if dtype is not None else out


def Assert(condition, data, **_kw):
    if not _np.all(condition):
        raise AssertionError("tf.Assert failed: %r" % (data,))


def assert_equal(x, y, **_kw):
    if not _np.all(_np.asarray(x) == _np.asarray(y)):
        raise AssertionError("tf.assert_equal failed")


def print(*args, **_kw):  # noqa: A001
    pass


# ---------------------------------------------------------------- constructors / shape ops
def convert_to_tensor(x, dtype=None, **_kw):
    return _t(x, dtype)


def constant(x, dtype=None, **_kw):
    return _t(x, dtype)


def cast(x, dtype):
    with _np.errstate(invalid="ignore"):
        return _np.asarray(x).astype(dtype)


def shape(x):
    return _np.asarray(_np.shape(x), dtype=_np.int32)


def reshape(x, shp):
    return _np.reshape(x, [int(s) for s in _np.asarray(shp).reshape(-1)])


def zeros(shp, dtype=float32):
    return _np.zeros([int(s) for s in _np.asarray(shp).reshape(-1)], dtype=dtype)


def ones(shp, dtype=float32):
    return _np.ones([int(s) for s in _np.asarray(shp).reshape(-1)], dtype=dtype)


def zeros_like(x):
    return _np.zeros_like(_np.asarray(x))


def ones_like(x):
    return _np.ones_like(_np.asarray(x))


def fill(dims, value):
    v = _np.asarray(value)
    return _np.full([int(s) for s in _np.asarray(dims).reshape(-1)], v, dtype=v.dtype)


def eye(n, batch_shape=None, dtype=float32):
    e = _np.eye(int(n), dtype=dtype)
    if batch_shape is not None:
        e = _np.broadcast_to(e, [int(s) for s in batch_shape] + [int(n), int(n)]).copy()
    return e


def range(*args, **_kw):  # noqa: A001
    return _np.arange(*[int(a) for a in args], dtype=_np.int32)


def meshgrid(*args):
    return _np.meshgrid(*args)


def expand_dims(x, axis):
    return _np.expand_dims(_np.asarray(x), int(axis))


def squeeze(x, axis=None):
    return _np.squeeze(_np.asarray(x), axis=axis)


def transpose(x, perm=None):
    return _np.transpose(x, perm)


def reverse(x, axis):
    return _np.flip(x, axis=tuple(int(a) for a in axis))


def stack(values, axis=0):
    return _np.stack(_pack(list(values)), axis=axis)


def concat(values, axis):
    return _np.concatenate(_pack(list(values)), axis=axis)


def tile(x, multiples):
    return _np.tile(x, [int(m) for m in multiples])


def broadcast_to(x, shp):
    return _np.broadcast_to(x, [int(s) for s in shp])


def one_hot(indices, depth, dtype=float32):
    return (_np.asarray(indices)[..., None] == _np.arange(int(depth))).astype(dtype)


def gather(params, indices, batch_dims=0, axis=None):
    params = _np.asarray(params)
    indices = _np.asarray(indices)
    if batch_dims == 0:
        return params[indices] if axis in (None, 0) else _np.take(params, indices, axis=axis)
    assert indices.ndim == batch_dims, "stand-in: gather(batch_dims=k) only with k-dim indices"
    grids = _np.meshgrid(*[_np.arange(s) for s in indices.shape], indexing="ij")
    return params[tuple(grids) + (indices,)]


def gather_nd(params, indices):
    indices = _np.asarray(indices)
    return _np.asarray(params)[tuple(_np.moveaxis(indices, -1, 0))]


def boolean_mask(tensor, mask):
    return _np.asarray(tensor)[_np.asarray(mask, dtype=_np.bool_)]


def where(condition, x=None, y=None):
    if x is None and y is None:
        return _np.argwhere(condition).astype(_np.int64)
    x = _like(x, y)
    y = _like(y, x)
    return _np.where(condition, x, y)


# ---------------------------------------------------------------- comparisons / logic
def _cmp(fn):
    def op(a, b):
        a = _like(a, b)
        b = _like(b, a)
        return fn(a, b)

    return op


less = _cmp(_np.less)
greater = _cmp(_np.greater)
equal = _cmp(_np.equal)
not_equal = _cmp(_np.not_equal)


def reduce_all(x, axis=None):
    return _np.all(x, axis=axis)


def reduce_any(x, axis=None):
    return _np.any(x, axis=axis)


# ---------------------------------------------------------------- arithmetic
def abs(x):  # noqa: A001
    return _np.abs(x)


def sqrt(x):
    with _np.errstate(invalid="ignore"):
        return _np.sqrt(x)


def square(x):
    return _np.multiply(x, x)


def sin(x):
    return _np.sin(x)


def cos(x):
    return _np.cos(x)


def reduce_sum(x, axis=None, keepdims=False):
    return _keepint(x, _np.sum, axis, keepdims)


def reduce_mean(x, axis=None):
    return _keepint(x, _np.mean, axis)


def reduce_min(x, axis=None):
    return _np.min(x, axis=axis)


def reduce_max(x, axis=None):
    return _np.max(x, axis=axis)


def argmax(x, axis=None, output_type=int64):
    x = _np.asarray(x)
    taps.append(("argmax", tuple(_map_index), x.copy()))
    return _np.argmax(x, axis=0 if axis is None else axis).astype(output_type)  # first maximum


def norm(x, ord="euclidean", axis=None, keepdims=False):  # noqa: A002
    # tf.norm(ord='euclidean') is sqrt(reduce_sum(x * x)) in the tensor's dtype
    assert ord in ("euclidean", 2)
    x = _np.asarray(x)
    with _np.errstate(over="ignore", invalid="ignore"):
        return _np.sqrt(_np.sum(x * x, axis=axis, keepdims=keepdims))


def matmul(a, b, transpose_a=False, transpose_b=False):
    a, b = _np.asarray(a), _np.asarray(b)
    if transpose_a:
        a = _np.swapaxes(a, -1, -2)
    if transpose_b:
        b = _np.swapaxes(b, -1, -2)
    with _np.errstate(over="ignore", invalid="ignore"):
        if a.dtype == _np.float32 and b.dtype == _np.float32 and ACCUMULATE == "float64":
            return _np.matmul(a.astype(_np.float64), b.astype(_np.float64)).astype(_np.float32)
        return _np.matmul(a, b)


class _Namespace:
    pass


math = _Namespace()
math.abs = abs
math.sqrt = sqrt
math.sin = sin
math.cos = cos
math.argmax = argmax
math.logical_and = _np.logical_and
math.is_nan = _np.isnan
math.is_finite = _np.isfinite
math.multiply = lambda a, b: _np.multiply(_like(a, b), _like(b, a))


def _divide(a, b):
    with _np.errstate(divide="ignore", invalid="ignore"):
        return _np.divide(_like(a, b), _like(b, a))


def _divide_no_nan(a, b):
    a, b = _np.asarray(_like(a, b)), _np.asarray(_like(b, a))
    with _np.errstate(divide="ignore", invalid="ignore"):
        q = _np.divide(a, b)
    return _np.where(b == 0, _np.zeros_like(q), q)


def _multiply_no_nan(x, y):
    x, y = _np.asarray(x), _np.asarray(y)
    with _np.errstate(invalid="ignore", over="ignore"):
        p = _np.multiply(x, y)
    return _np.where(y == 0, _np.zeros_like(p), p)


def _softplus(x):
    # TensorFlow's functor (core/kernels/softplus_op.h): x above -threshold passes through, below
    # threshold is exp(x), otherwise log1p(exp(x)); threshold = log(eps) + 2
    x = _np.asarray(x)
    thr = _np.log(_np.finfo(x.dtype).eps).astype(x.dtype) + x.dtype.type(2)
    with _np.errstate(over="ignore"):
        e = _np.exp(x)
    return _np.where(x > -thr, x, _np.where(x < thr, e, _np.log1p(e)))


def _count_nonzero(x, axis=None):
    return _np.count_nonzero(x, axis=axis).astype(_np.int64) if axis is not None else _np.int64(_np.count_nonzero(x))


def _bincount(arr, minlength=None, axis=None, **_kw):
    arr = _np.asarray(arr)
    if arr.ndim == 1:
        return _np.bincount(arr, minlength=minlength or 0).astype(arr.dtype)
    assert axis in (-1, arr.ndim - 1) and arr.ndim == 2
    size = _bi.max(int(arr.max()) + 1 if arr.size else 0, int(minlength or 0))
    out = _np.zeros((arr.shape[0], size), dtype=arr.dtype)
    for r in _bi.range(arr.shape[0]):
        out[r] = _np.bincount(arr[r], minlength=size)
    return out


def _top_k(x, k=1, **_kw):
    # descending, ties broken towards the lower index (TensorFlow's documented rule)
    x = _np.asarray(x)
    idx = _np.argsort(-x.astype(_np.int64) if x.dtype.kind in "iu" else -x, axis=-1, kind="stable")[..., : int(k)]
    return _np.take_along_axis(x, idx, axis=-1), idx.astype(_np.int32)


math.divide = _divide
math.divide_no_nan = _divide_no_nan
math.multiply_no_nan = _multiply_no_nan
math.softplus = _softplus
math.count_nonzero = _count_nonzero
math.bincount = _bincount
math.top_k = _top_k


def _softmax(x, axis=-1):
    x = _np.asarray(x)
    with _np.errstate(over="ignore", invalid="ignore"):
        e = _np.exp(x - _np.max(x, axis=axis, keepdims=True))
        return e / _np.sum(e, axis=axis, keepdims=True)


nn = _Namespace()
nn.softmax = _softmax
nn.sigmoid = lambda x: (1 / (1 + _np.exp(-_np.asarray(x)))).astype(_np.asarray(x).dtype)
nn.relu = lambda x: _np.maximum(x, _np.zeros((), dtype=_np.asarray(x).dtype))


def _pinv(a, rcond=None):
    a = _np.asarray(a)
    if rcond is None:  # tf.linalg.pinv default: 10 * max(rows, cols) * eps
        rcond = 10.0 * _bi.max(a.shape[-2:]) * _np.finfo(a.dtype).eps
    return _np.linalg.pinv(a, rcond=rcond)


def _svd(x, compute_uv=True, **_kw):
    assert not compute_uv
    return _np.linalg.svd(_np.asarray(x), compute_uv=False)


linalg = _Namespace()
linalg.svd = _svd
linalg.inv = lambda x: _np.linalg.inv(_np.asarray(x))
linalg.pinv = _pinv
linalg.matmul = matmul
linalg.tensor_diag_part = lambda x: _np.diagonal(x)


class _Random:
    provider = None

    def uniform(self, shp, minval=0, maxval=None, dtype=float32, **_kw):
        if self.provider is None:
            raise RuntimeError("tf stand-in: set tf.random.provider before a graph that draws random numbers")
        out = self.provider(tuple(int(s) for s in shp), minval, maxval, dtype, tuple(_map_index))
        return _np.asarray(out).astype(dtype)


random = _Random()


# ---------------------------------------------------------------- tf.keras.layers.Layer
class _Layer:
    def __init__(self, name=None, **_kw):
        self.name = name
        self._built = False

    def build(self, input_shape):
        pass

    def __call__(self, inputs, **kwargs):
        if not self._built:
            if isinstance(inputs, (list, tuple)):
                self.build([tuple(_np.shape(i)) for i in inputs])
            else:
                self.build(tuple(_np.shape(inputs)))
            self._built = True
        return self.call(inputs, **kwargs)


keras = _Namespace()
keras.layers = _Namespace()
keras.layers.Layer = _Layer
