"""Philox4x32-10 counter-based RNG (numpy) and the two random streams of the voting path.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference draws its random numbers from ``tf.random.uniform``
(/root/reference/casapose/pose_estimation/ransac_voting.py:296 ``selection`` and
:319-321 ``idxs``).  TensorFlow's stream cannot be reproduced here (TF is not
installed), so both the oracle and the CUDA path consume the SAME explicit stream,
defined in this file and restated in casapose_b200/csrc/philox.cuh:

  key      = (seed & 0xffffffff, seed >> 32)
  idxs     : element e = (h*vn + v)*2 + k   of job (image, cls) in round r
             counter = (e >> 2, r, cls | STREAM_IDXS << 16, image), word = out[e & 3]
             idx     = (word * tn) >> 32                      (uniform in [0, tn))
  selection: element e = y*w + x             of job (image, cls)
             counter = (e >> 2, 0, cls | STREAM_SELECTION << 16, image), word = out[e & 3]
             u       = float32(word >> 8) * 2**-24            (uniform in [0, 1))

The generator itself is the published Philox4x32-10 of Salmon et al. (SC'11,
"Parallel random numbers: as easy as 1, 2, 3"); `philox4x32_10` is checked against the
Random123 known-answer vectors in tests/test_philox.py.
"""
import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

STREAM_IDXS = 0
STREAM_SELECTION = 1


def philox4x32_10(counter, key):
    """counter: uint32 array [..., 4]; key: (k0, k1) python ints. Returns uint32 [..., 4]."""
    counter = np.asarray(counter, dtype=np.uint32)
    c0 = counter[..., 0].astype(np.uint64)
    c1 = counter[..., 1].astype(np.uint64)
    c2 = counter[..., 2].astype(np.uint64)
    c3 = counter[..., 3].astype(np.uint64)
    k0 = int(key[0]) & 0xFFFFFFFF
    k1 = int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = PHILOX_M0 * c0  # 32x32 -> 64 bit, exact in uint64
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def _key(seed):
    seed = int(seed)
    return (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)


def _words(n_elems, c1, c2, c3, seed):
    """First n_elems 32-bit words of the stream with fixed (c1, c2, c3)."""
    n_blocks = (n_elems + 3) // 4
    ctr = np.zeros((n_blocks, 4), dtype=np.uint32)
    ctr[:, 0] = np.arange(n_blocks, dtype=np.uint32)
    ctr[:, 1] = c1
    ctr[:, 2] = c2
    ctr[:, 3] = c3
    return philox4x32_10(ctr, _key(seed)).reshape(-1)[:n_elems]


def raw_idx_words(seed, image, cls, rnd, hn, vn):
    """uint32 [hn, vn, 2] raw words for one job and round (before reduction to [0, tn))."""
    w = _words(hn * vn * 2, rnd, (cls & 0xFFFF) | (STREAM_IDXS << 16), image, seed)
    return w.reshape(hn, vn, 2)


def draw_idxs(seed, image, cls, rnd, hn, vn, tn):
    """int32 [hn, vn, 2] pixel-pair indices in [0, tn) — stands in for ransac_voting.py:319-321."""
    w = raw_idx_words(seed, image, cls, rnd, hn, vn).astype(np.uint64)
    return ((w * np.uint64(tn)) >> np.uint64(32)).astype(np.int32)


def draw_selection(seed, image, cls, h, w):
    """float32 [h, w] uniform [0,1) — stands in for ransac_voting.py:296."""
    words = _words(h * w, 0, (cls & 0xFFFF) | (STREAM_SELECTION << 16), image, seed)
    u = (words >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    return u.reshape(h, w)
