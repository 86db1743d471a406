"""torch stand-in for the TensorFlow calls of the reference's CoordLSVotingWeighted — used ONLY to obtain the
layer's gradient from the reference's own forward code.

TEST INFRASTRUCTURE (see oracle/__init__.py, oracle/tf_standin/tensorflow/__init__.py).  The reference trains
through this layer with TensorFlow's autodiff (train_casapose.py:536-595).  No TensorFlow can run here, so
oracle/make_golden.py executes /root/reference/casapose/pose_estimation/voting_layers_2d.py UNMODIFIED with this
module as ``tf``: every op maps to the torch op of the same meaning in the same dtype (float32 elementwise,
float64 sums and pinv, as the reference casts them), and torch.autograd differentiates the graph the reference's
code built.  Both autodiff systems return the derivative of the same composition wherever it exists; the inputs
of the golden case avoid the one place they differ (sqrt at an exactly zero vector: NaN in TensorFlow)."""
import builtins as _bi

import numpy as _np
import torch as _torch

__version__ = "0.0-torch-standin"

float32 = _torch.float32
float64 = _torch.float64
int32 = _torch.int32
int64 = _torch.int64


def _dt(d):
    return d if isinstance(d, _torch.dtype) else {_np.dtype("float32"): float32, _np.dtype("float64"): float64,
                                                  _np.dtype("int32"): int32}[_np.dtype(d)]


def _t(x, like=None):
    if isinstance(x, _torch.Tensor):
        return x
    if like is not None:
        return _torch.as_tensor(x, dtype=like.dtype)
    a = _torch.as_tensor(x)
    if a.dtype == _torch.float64:
        return a.to(float32)
    if a.dtype == _torch.int64:
        return a.to(int32)
    return a


def function(fn=None, **_kw):
    return fn if fn is not None else (lambda f: f)


def stop_gradient(x):
    return x.detach()


def cast(x, dtype):
    return _t(x).to(_dt(dtype)) if isinstance(x, _torch.Tensor) else _torch.as_tensor(x, dtype=_dt(dtype))


def constant(x, dtype=None):
    return _torch.as_tensor(x, dtype=_dt(dtype)) if dtype is not None else _t(x)


def expand_dims(x, axis):
    return _torch.unsqueeze(x, int(axis))


def squeeze(x, axis=None):
    return _torch.squeeze(x) if axis is None else _torch.squeeze(x, int(axis))


def reshape(x, shp):
    return _torch.reshape(x, [int(s) for s in shp])


def transpose(x, perm=None):
    return x.permute(*perm) if perm is not None else x.t()


def shape(x):
    return list(x.shape)


def eye(n, dtype=float32):
    return _torch.eye(int(n), dtype=dtype)


def range(n):  # noqa: A001
    return _torch.arange(int(n), dtype=int32)


def meshgrid(a, b):
    g = _torch.meshgrid(a, b, indexing="xy")
    return g[0], g[1]


def stack(values, axis=0):
    return _torch.stack(list(values), dim=axis)


def where(cond, x, y):
    x = x if isinstance(x, _torch.Tensor) else _t(x, y if isinstance(y, _torch.Tensor) else None)
    y = y if isinstance(y, _torch.Tensor) else _t(y, x)
    return _torch.where(cond, x, y)


def map_fn(fn, elems, dtype=None, **_kw):
    return _torch.stack([fn(elems[i]) for i in _bi.range(len(elems))], dim=0)


def norm(x, ord="euclidean", axis=None, keepdims=False):  # noqa: A002
    return _torch.sqrt(_torch.sum(x * x, dim=axis, keepdim=keepdims))


def matmul(a, b):
    return _torch.matmul(a, b)


def reduce_sum(x, axis=None, keepdims=False):
    return _torch.sum(x, dim=tuple(axis) if isinstance(axis, (list, tuple)) else axis, keepdim=keepdims)


def reduce_all(x):
    return _torch.all(x)


def Assert(condition, data, **_kw):
    if not _bi.bool(condition):
        raise AssertionError("tf.Assert failed: %r" % (data,))


class _NS:
    pass


def _divide_no_nan(x, y):
    safe = _torch.where(y == 0, _torch.ones_like(y), y)
    return _torch.where(y == 0, _torch.zeros_like(x / safe), x / safe)


def _multiply_no_nan(x, y):
    return _torch.where(y == 0, _torch.zeros_like(x * y), x * y)


def _bincount(arr, minlength=None, axis=None, **_kw):
    a = arr.numpy()
    size = _bi.max(int(a.max()) + 1, int(minlength or 0))
    return _torch.as_tensor(_np.stack([_np.bincount(r, minlength=size) for r in a]).astype(_np.int32))


def _top_k(x, k=1, **_kw):
    a = x.numpy().astype(_np.int64)
    idx = _np.argsort(-a, axis=-1, kind="stable")[..., : int(k)]
    return _torch.as_tensor(_np.take_along_axis(a, idx, -1).astype(_np.int32)), _torch.as_tensor(idx.astype(_np.int32))


math = _NS()
math.softplus = _torch.nn.functional.softplus
math.divide_no_nan = _divide_no_nan
math.multiply_no_nan = _multiply_no_nan
math.is_finite = _torch.isfinite
math.bincount = _bincount
math.top_k = _top_k

nn = _NS()
nn.sigmoid = _torch.sigmoid
nn.softmax = lambda x: _torch.softmax(x, dim=-1)

linalg = _NS()
linalg.pinv = lambda a: _torch.linalg.pinv(a, rtol=10.0 * _bi.max(a.shape[-2:]) * _torch.finfo(a.dtype).eps)


class _Layer:
    def __init__(self, name=None, **_kw):
        self.name = name
        self._built = False

    def __call__(self, inputs, **kwargs):
        if not self._built:
            self.build([tuple(i.shape) for i in inputs])
            self._built = True
        return self.call(inputs, **kwargs)


keras = _NS()
keras.layers = _NS()
keras.layers.Layer = _Layer
