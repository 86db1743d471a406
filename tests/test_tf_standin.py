"""The numpy TensorFlow stand-in (oracle/tf_standin) is what lets the reference's own source produce the golden
vectors, so its semantics are tested against TensorFlow's DOCUMENTED behaviour on hand-written cases: dtype
defaults, tie rules, ordering, immutability.  (No TensorFlow is importable here)"""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _load(name, *parts):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "oracle", *parts, "__init__.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def tf():
    return _load("_standin_tf", "tf_standin", "tensorflow")


@pytest.fixture(scope="module")
def tfa():
    return _load("_standin_tfa", "tf_standin", "tensorflow_addons")


def test_python_scalars_become_float32_and_int32(tf):
    assert tf.convert_to_tensor(0.99).dtype == np.float32 and tf.convert_to_tensor(20).dtype == np.int32
    assert tf.constant([1, -1], dtype=tf.float32).dtype == np.float32
    x = np.ones(3, np.float32)
    assert tf.where(x > 0, x, [0, 0, 0]).dtype == np.float32  # a python list takes the tensor's dtype
    assert tf.where(x > 0, 1, 0).dtype == np.int32
    assert tf.stack([1.0, 0.0, -x[0]]).dtype == np.float32  # auto-packing follows the tensor in the list
    assert tf.greater(x, 1e-6).dtype == np.bool_ and (x * 0.1).dtype == np.float32
    # float32 threshold, not float64: 0.99f < 0.99
    a = np.float32(0.99) + np.float32(2.0 ** -25)
    assert bool(tf.greater(np.float32(a), tf.convert_to_tensor(0.99))) == bool(np.float32(a) > np.float32(0.99))


def test_integer_reductions_keep_their_dtype(tf):
    v = np.ones((4, 5), np.int32)
    assert tf.reduce_sum(v, 1).dtype == np.int32 and tf.reduce_max(v, 0).dtype == np.int32
    assert tf.math.count_nonzero(v, 1).dtype == np.int64 and tf.argmax(v, 0).dtype == np.int64
    assert tf.shape(v).dtype == np.int32 and tf.range(3).dtype == np.int32


def test_tie_rules(tf):
    c = np.array([[3, 7, 7], [7, 7, 1]], np.int32)
    assert tf.argmax(c, 0).tolist() == [1, 0, 0]  # first maximum
    vals, idx = tf.math.top_k(np.array([[5, 0, 0, 9, 5]], np.int32), k=3)
    assert vals.tolist() == [[9, 5, 5]] and idx.tolist() == [[3, 0, 4]]  # equal elements: lower index first
    _, idx = tf.math.top_k(np.array([[100, 0, 0]], np.int32), k=2)
    assert idx.tolist() == [[0, 1]]  # all-zero tail: label 1 (voting_layers_2d.py:67-75 relies on this)


def test_where_boolean_mask_gather_orders(tf):
    m = np.array([[0, 1, 0], [1, 1, 0]], np.float32)
    assert tf.where(tf.not_equal(m, 0.0)).tolist() == [[0, 1], [1, 0], [1, 1]]  # row-major coordinates (y, x)
    v = np.arange(12, dtype=np.float32).reshape(2, 3, 2)
    assert tf.boolean_mask(v, tf.cast(m, tf.bool)).tolist() == [[2, 3], [6, 7], [8, 9]]
    assert tf.reverse(v, axis=[2])[0, 0].tolist() == [1, 0]
    params = np.arange(24).reshape(2, 3, 4)
    idx = np.array([[1, 2], [0, 0]])
    assert tf.gather_nd(params, idx).tolist() == [params[1, 2].tolist(), params[0, 0].tolist()]
    p6 = np.arange(2 * 2 * 2 * 3 * 2).reshape(2, 2, 2, 3, 2)
    sel = np.array([[[0, 2], [1, 1]], [[2, 0], [0, 1]]])
    out = tf.gather(p6, sel, batch_dims=3)
    assert out.shape == (2, 2, 2, 2) and out[1, 0, 0].tolist() == p6[1, 0, 0, 2].tolist()
    assert tf.one_hot(np.array([2, 0]), 3).tolist() == [[0, 0, 1], [1, 0, 0]]


def test_bincount_rows_share_one_length(tf):
    comp = np.array([[0, 0, 1, 1, 1], [0, 3, 3, 0, 0]], np.int32)
    b = tf.math.bincount(comp, axis=-1, minlength=2)
    assert b.tolist() == [[2, 3, 0, 0], [3, 0, 0, 2]] and b.dtype == np.int32


def test_map_fn_hands_out_copies_and_tracks_indices(tf):
    x = np.ones((2, 3), np.float32)
    seen = []

    def body(row):
        seen.append(tuple(tf._map_index))
        row *= 0  # TensorFlow tensors are immutable: this must not reach the caller's array
        return row

    out = tf.map_fn(lambda r: tf.map_fn(lambda e: body(e), r, dtype=tf.float32), x, dtype=tf.float32)
    assert x.all() and not out.any() and seen[:4] == [(0, 0), (0, 1), (0, 2), (1, 0)]


def test_no_nan_ops_softmax_softplus_pinv(tf):
    assert tf.math.divide_no_nan(np.float32([1, 2]), np.float32([0, 4])).tolist() == [0.0, 0.5]
    assert tf.math.multiply_no_nan(np.float32([np.inf, 2]), np.float32([0, 3])).tolist() == [0.0, 6.0]
    hot = tf.nn.softmax(np.float32([[1.0, 1.0 + 1e-3, 0.0]]) * np.float32(1e6))
    assert hot.dtype == np.float32 and hot.tolist() == [[0.0, 1.0, 0.0]]
    tie = tf.nn.softmax(np.float32([[2.0, 2.0, 0.0]]) * np.float32(1e6))
    assert tie.tolist() == [[0.5, 0.5, 0.0]] and tf.cast(tie + 0.1, tf.int32).tolist() == [[0, 0, 0]]  # :44 drops exact ties
    x = np.float32([-30.0, 0.0, 30.0])
    assert np.allclose(tf.math.softplus(x), np.log1p(np.exp(x.astype(np.float64))), rtol=1e-6)
    r = np.array([[4.0, 0.0], [0.0, 1e-20]])
    assert np.allclose(tf.linalg.pinv(r), [[0.25, 0.0], [0.0, 0.0]])  # rcond = 10 * 2 * eps cuts the tiny value


def test_float32_reductions_round_once(tf):
    v = (np.float32(1.0) + np.arange(4096, dtype=np.float32) * np.float32(2.0 ** -20)).astype(np.float32)
    exact = np.float32(v.astype(np.float64).sum())
    assert tf.reduce_sum(v) == exact
    a = np.random.default_rng(0).normal(size=(2, 5000)).astype(np.float32)
    assert np.array_equal(tf.matmul(a, a.T), (a.astype(np.float64) @ a.astype(np.float64).T).astype(np.float32))


def test_connected_components_contract(tfa):
    img = np.array([[1, 1, 0, 1],
                    [0, 0, 0, 1],
                    [1, 0, 1, 0],
                    [1, 0, 1, 1]], np.int32)
    lab = tfa.image.connected_components(img)
    # 4-connectivity (the diagonal neighbours (1,3)/(2,2) stay apart), ids in row-major order of the first pixel
    assert lab.tolist() == [[1, 1, 0, 2], [0, 0, 0, 2], [3, 0, 4, 0], [3, 0, 4, 4]] and lab.dtype == np.int32
    two = np.array([[1, 2, 2], [1, 0, 2]], np.int32)  # different values never join
    assert tfa.image.connected_components(two).tolist() == [[1, 2, 2], [1, 0, 2]]
