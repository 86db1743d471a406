"""Philox4x32-10 against the published Random123 known-answer vectors, and the stream definitions."""
import numpy as np

from oracle import philox_np as P

# Random123 kat_vectors: philox4x32 10 rounds (counter, key, expected)
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF), (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def test_known_answer_vectors():
    for ctr, key, exp in KAT:
        got = P.philox4x32_10(np.array(ctr, dtype=np.uint32), key)
        assert tuple(int(x) for x in got) == exp


def test_vectorised_matches_scalar():
    ctr = np.arange(40, dtype=np.uint32).reshape(10, 4)
    got = P.philox4x32_10(ctr, (5, 6))
    for i in range(10):
        assert np.array_equal(got[i], P.philox4x32_10(ctr[i], (5, 6)))


def test_idx_stream_range_and_determinism():
    a = P.draw_idxs(1237, 3, 2, 1, 64, 9, 777)
    b = P.draw_idxs(1237, 3, 2, 1, 64, 9, 777)
    assert a.shape == (64, 9, 2) and a.dtype == np.int32
    assert np.array_equal(a, b)
    assert a.min() >= 0 and a.max() < 777
    assert not np.array_equal(a, P.draw_idxs(1237, 3, 2, 2, 64, 9, 777))  # other round
    assert not np.array_equal(a, P.draw_idxs(1237, 4, 2, 1, 64, 9, 777))  # other image
    # raw words do not depend on tn; the reduction is (word * tn) >> 32
    w = P.raw_idx_words(1237, 3, 2, 1, 64, 9).astype(np.uint64)
    assert np.array_equal(a, ((w * np.uint64(777)) >> np.uint64(32)).astype(np.int32))


def test_selection_stream():
    s = P.draw_selection(9, 0, 1, 13, 17)
    assert s.shape == (13, 17) and s.dtype == np.float32
    assert s.min() >= 0.0 and s.max() < 1.0
    assert abs(float(P.draw_selection(9, 0, 1, 200, 200).mean()) - 0.5) < 0.01
