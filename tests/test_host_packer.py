"""The host mask packer of casa_ransac_vote_host (AVX2 / scalar paths, chunked work distribution over threads and
ranges) against numpy: bit c of a pixel's word = (mask[p][c] != 0) as tf.not_equal counts it
(/root/reference/casapose/pose_estimation/ransac_voting.py:304).  No device involved."""
import ctypes as C

import numpy as np
import pytest


def _want(mask):
    oc = mask.shape[1]
    with np.errstate(invalid="ignore"):
        set_ = mask != 0  # NaN != 0 is True, -0.0 != 0 is False
    bits = (set_.astype(np.uint32) << np.arange(oc, dtype=np.uint32)).sum(axis=1, dtype=np.uint32)
    not_binary = bool((set_ & ~(mask == 1.0)).any())
    return bits, not_binary


def _pack(lib, mask, threads, parts):
    mask = np.ascontiguousarray(mask, np.float32)
    bits = np.full(mask.shape[0], 0xDEADBEEF, np.uint32)
    rc = lib.casa_selftest_pack(mask.ctypes.data, bits.ctypes.data, mask.shape[0], mask.shape[1], threads, parts)
    assert rc in (0, 1), lib.casa_last_error()
    return bits, bool(rc)


@pytest.mark.parametrize("oc", [8, 3, 13, 32])
@pytest.mark.parametrize("threads,parts", [(1, 1), (4, 4), (7, 3)])
def test_one_hot_masks(cuda_lib, oc, threads, parts):
    rng = np.random.default_rng(oc * 100 + threads)
    npx = 100003  # not a multiple of the chunk, of 8, or of the thread count
    labels = rng.integers(0, oc + 4, size=npx)  # values above oc - 1: background rows (all zero)
    mask = np.zeros((npx, oc), np.float32)
    rows = np.nonzero(labels < oc)[0]
    mask[rows, labels[rows]] = 1.0
    got, nb = _pack(cuda_lib, mask, threads, parts)
    want, wnb = _want(mask)
    assert np.array_equal(got, want) and nb == wnb and not nb


@pytest.mark.parametrize("oc", [8, 5])
def test_special_values_and_overlaps(cuda_lib, oc):
    rng = np.random.default_rng(7)
    npx = 40000
    mask = (rng.uniform(size=(npx, oc)) < 0.2).astype(np.float32)  # overlapping channels
    got, nb = _pack(cuda_lib, mask, 5, 2)
    want, wnb = _want(mask)
    assert np.array_equal(got, want) and nb == wnb and not nb
    mask[11, 0] = -0.0   # not set
    mask[12, 1] = np.nan  # set, not binary
    mask[13, 2] = 0.5     # set, not binary
    mask[14, 3] = -1.0
    mask[15, oc - 1] = np.float32(1e-45)  # denormal: non-zero
    got, nb = _pack(cuda_lib, mask, 5, 2)
    want, wnb = _want(mask)
    assert np.array_equal(got, want) and nb and wnb
    assert not (got[11] & 1) or mask[11, 0] != 0


def test_empty_and_tiny_inputs(cuda_lib):
    mask = np.zeros((0, 8), np.float32)
    bits = np.zeros(1, np.uint32)
    assert cuda_lib.casa_selftest_pack(mask.ctypes.data if mask.size else bits.ctypes.data, bits.ctypes.data, 0, 8, 3, 2) == 0
    for npx in (1, 7, 9):
        m = np.ones((npx, 8), np.float32)
        got, nb = _pack(cuda_lib, m, 3, 2)
        assert np.array_equal(got, np.full(npx, 0xFF, np.uint32)) and not nb
    assert cuda_lib.casa_selftest_pack(None, bits.ctypes.data, 1, 8, 1, 1) == -1
    assert cuda_lib.casa_selftest_pack(bits.ctypes.data, bits.ctypes.data, 1, 40, 1, 1) == -1
