"""Pipelined host entry (casa_ransac_vote_host_async / casa_host_wait): calls in flight return exactly what the
synchronous host call and the device call return, in any waiting order; errors surface at the wait; the golden vectors
of the reference's code hold through it."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from .test_golden_oracle import load, ransac_case_inputs  # noqa: E402

pytestmark = pytest.mark.gpu
TOL_PX = 1e-3


def _frames(b, seed, variant="easy", h=240, w=320):
    from casapose_b200 import synthetic

    d = synthetic.make_frames(b, h, w, synthetic.CONFIG_8_IDS, seed=seed, variant=variant)
    return torch.from_numpy(d["mask"]).pin_memory(), torch.from_numpy(d["vertex"]).pin_memory()


def test_calls_in_flight_equal_synchronous_calls(cuda_lib):
    from casapose_b200.pose_estimation.ransac_voting import (ransac_voting_layer_all_masks,
                                                               ransac_voting_layer_all_masks_host)

    inputs = [_frames(8, 11), _frames(8, 12, "hard"), _frames(4, 13), _frames(8, 14), _frames(1, 15), _frames(8, 16)]
    want = [ransac_voting_layer_all_masks_host(m, v, 128, seed=40 + i).clone() for i, (m, v) in enumerate(inputs)]
    dev = [ransac_voting_layer_all_masks(m.cuda(), v.cuda(), 128, seed=40 + i).cpu() for i, (m, v) in enumerate(inputs)]
    for a, b in zip(want, dev):
        assert torch.equal(a, b)
    for rep in range(3):  # every call issued before the first wait; the third and later ones block for a free driver
        pend = [ransac_voting_layer_all_masks_host(m, v, 128, seed=40 + i, wait=False) for i, (m, v) in enumerate(inputs)]
        order = range(len(pend)) if rep != 1 else reversed(range(len(pend)))
        for i in order:
            assert torch.equal(pend[i].result(), want[i]), "call %d differs (pass %d)" % (i, rep)
        assert torch.equal(pend[0].result(), want[0])  # a second result() is a no-op


def test_pipelined_calls_hold_the_golden_vectors(cuda_lib):
    from casapose_b200.pose_estimation.ransac_voting import ransac_voting_layer_all_masks_host

    cases = []
    for name in ("ransac_full_480x640", "ransac_hard", "ransac_full_480x640"):
        g = load(name)
        mask, vertex, hn, seed, kw = ransac_case_inputs(g)
        tm = torch.from_numpy(np.ascontiguousarray(mask)).pin_memory()
        tv = torch.from_numpy(np.ascontiguousarray(vertex)).pin_memory()
        cases.append((g["points"], ransac_voting_layer_all_masks_host(tm, tv, hn, seed=seed, wait=False, **kw)))
    for points, pend in cases:
        assert np.abs(pend.result().numpy() - points).max() <= TOL_PX


def test_synchronous_and_pipelined_calls_mix(cuda_lib):
    from casapose_b200.pose_estimation.ransac_voting import ransac_voting_layer_all_masks_host

    m, v = _frames(8, 21)
    want = ransac_voting_layer_all_masks_host(m, v, 128, seed=3).clone()
    p0 = ransac_voting_layer_all_masks_host(m, v, 128, seed=3, wait=False)
    got_sync = ransac_voting_layer_all_masks_host(m, v, 128, seed=3).clone()  # the handle's own call beside a driver's
    p1 = ransac_voting_layer_all_masks_host(m, v, 128, seed=3, wait=False)
    assert torch.equal(got_sync, want) and torch.equal(p0.result(), want) and torch.equal(p1.result(), want)


def test_errors_surface_at_submit_or_wait(cuda_lib):
    from casapose_b200 import _lib
    from casapose_b200.pose_estimation.ransac_voting import _params

    lib = cuda_lib
    hdl = _lib.handle(0)
    m, v = _frames(2, 31)
    out = torch.empty((2, 8, 9, 2), dtype=torch.float32).pin_memory()
    ticket = C.c_int64(-1)
    bad = _params(2, 240, 320, 40, 9, 128, 0.99, 0.99, 20, 5, 30000, 0, 0, 0, False)  # oc = 40: rejected at once
    assert lib.casa_ransac_vote_host_async(hdl, C.byref(bad), m.data_ptr(), v.data_ptr(), out.data_ptr(), C.byref(ticket)) != 0
    assert b"oc" in lib.casa_last_error()
    assert lib.casa_host_wait(hdl, 10 ** 9) != 0  # unknown ticket
    # overlapping channels overflow the default pixel lists: the error belongs to the call, reported by its wait
    both = torch.ones((2, 240, 320, 8), dtype=torch.float32).pin_memory()
    p = _params(2, 240, 320, 8, 9, 128, 0.99, 0.99, 20, 5, 30000, 0, 0, 0, False)
    assert lib.casa_ransac_vote_host_async(hdl, C.byref(p), both.data_ptr(), v.data_ptr(), out.data_ptr(), C.byref(ticket)) == 0
    rc = lib.casa_host_wait(hdl, ticket.value)
    assert rc == -3 and b"pix" in lib.casa_last_error().lower()
    # the handle works on
    good = _params(2, 240, 320, 8, 9, 128, 0.99, 0.99, 20, 5, 30000, 7, 0, 0, False)
    assert lib.casa_ransac_vote_host_async(hdl, C.byref(good), m.data_ptr(), v.data_ptr(), out.data_ptr(), C.byref(ticket)) == 0
    assert lib.casa_host_wait(hdl, ticket.value) == 0
    # a call nobody waits for keeps its error for casa_sync
    assert lib.casa_ransac_vote_host_async(hdl, C.byref(p), both.data_ptr(), v.data_ptr(), out.data_ptr(), C.byref(ticket)) == 0
    assert lib.casa_sync(hdl) == -3
    assert lib.casa_sync(hdl) == 0
