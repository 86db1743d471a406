"""Backward of CoordLSVotingWeighted (casa_ls_vote_backward).

CPU: the gradient oracle (torch float64 autograd over the restated `calc`) agrees with the numpy forward oracle
and with central finite differences.  GPU: the closed-form CUDA adjoint against the gradient oracle."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from casapose_b200 import synthetic  # noqa: E402
from oracle import ls_voting_grad as G  # noqa: E402
from oracle import ls_voting_np as L  # noqa: E402

F = np.float32


def _frames(b=2, h=120, w=160, ids=(1, 5, 6), seed=0):
    d = synthetic.make_frames(b, h, w, ids, variant="easy", with_logits=True)
    rng = np.random.default_rng(seed)
    g = rng.normal(size=(b, len(ids), 9, 2)).astype(F)
    return d["seg_logits"], d["vertex"].reshape(b, h, w, 18).copy(), d["conf_logits"], g


def _mask_singular(g, seg, direct, conf, keep_rank_one=(), **kw):
    """Zero the incoming gradient of every (image, class, keypoint) whose 2x2 system is numerically singular
    (a component of a few pixels, an empty class): the derivative of an ill-conditioned pseudo-inverse amplifies
    last-ulp differences by its condition number and is not a meaningful comparison."""
    _, dbg = L.coord_ls_voting_weighted(seg, direct, conf, return_debug=True, **kw)
    sv = np.linalg.svd(dbg["R"], compute_uv=False)  # [b,oc,vn,2]
    bad = ~(sv[..., 1] > 1e-4 * sv[..., 0])
    for k in keep_rank_one:
        bad[:, :, k] &= sv[:, :, k, 1] != 0  # exactly rank one on both sides: the constant-rank formula applies
    g = g.copy()
    g[bad] = 0
    return g


def test_gradient_oracle_forward_equals_numpy_oracle_and_finite_differences():
    seg, direct, conf, g = _frames(b=1, h=40, w=56, ids=(1, 5))
    out, gd, gw = G.ls_vote_with_grads(seg, direct, conf, g)
    ref = L.coord_ls_voting_weighted(seg, direct, conf)
    assert np.abs(out - ref).max() < 1e-3
    hot = L.hard_softmax_f32(seg)[..., 1:]
    ys, xs, cs = np.nonzero(hot[0])
    rng = np.random.default_rng(1)
    for k in rng.choice(len(ys), size=4, replace=False):
        y, x = ys[k], xs[k]
        ch = int(rng.integers(0, 18))
        eps = 1e-4
        dp, dm = direct.astype(np.float64), direct.astype(np.float64)
        dp[0, y, x, ch] += eps
        dm[0, y, x, ch] -= eps
        lp = (G.ls_vote_with_grads(seg, dp, conf, g)[0] * g).sum()
        lm = (G.ls_vote_with_grads(seg, dm, conf, g)[0] * g).sum()
        fd = (lp - lm) / (2 * eps)
        assert abs(fd - gd[0, y, x, ch]) <= 1e-5 * max(1.0, abs(fd)) + 1e-7 * np.abs(gd).max(), (fd, gd[0, y, x, ch])
    assert not gd[0][hot[0].sum(-1) == 0].any() and not gw[0][hot[0].sum(-1) == 0].any()  # no gradient off the masks


def _check(a, ref, what):
    scale = np.abs(ref).max()
    err = np.abs(a - ref).max()
    assert err <= 2e-4 * scale, "%s: max err %.3g of scale %.3g" % (what, err, scale)
    assert np.array_equal(a != 0, ref != 0) or np.abs(a[(a != 0) != (ref != 0)]).max() <= 1e-6 * scale


@pytest.mark.gpu
@pytest.mark.parametrize("filt", [False, True])
@pytest.mark.parametrize("sigm", [False, True])
def test_gpu_backward_matches_gradient_oracle(cuda_lib, filt, sigm):
    from casapose_b200.pose_estimation import CoordLSVotingWeighted

    seg, direct, conf, g = _frames()
    g = _mask_singular(g, seg, direct, conf, filter_estimates=filt, sigmoid_weights=sigm)
    assert (g != 0).any(axis=(2, 3)).sum() >= 4  # most jobs are regular
    layer = CoordLSVotingWeighted("ls", seg.shape[3], filter_estimates=filt, sigmoid_weights=sigm)
    inp = [torch.from_numpy(seg).cuda(), torch.from_numpy(direct).cuda(), torch.from_numpy(conf).cuda()]
    gd, gw = layer.backward(inp, torch.from_numpy(g).cuda())
    _, rd, rw = G.ls_vote_with_grads(seg, direct, conf, g, filter_estimates=filt, sigmoid_weights=sigm)
    _check(gd.cpu().numpy(), rd, "grad_direct")
    _check(gw.cpu().numpy(), rw, "grad_conf")


@pytest.mark.gpu
def test_gpu_backward_through_torch_autograd_and_degenerate_inputs(cuda_lib):
    from casapose_b200.pose_estimation import CoordLSVotingWeighted

    seg, direct, conf, g = _frames(b=1, h=48, w=64, ids=(1, 5))
    direct[0, :, :, 4:6] = np.array([1.0, 0.0], F)  # keypoint 2: every vector (dy,dx) = (1,0) -> rank-one system
    hot = L.hard_softmax_f32(seg)[..., 1:]
    ys, xs, _ = np.nonzero(hot[0])
    direct[0, ys[3], xs[3], 0:2] = 0.0  # a zero vector on a masked pixel: zero gradient, no NaN
    g = _mask_singular(g, seg, direct, conf, keep_rank_one=(2,))
    layer = CoordLSVotingWeighted("ls", seg.shape[3])
    ts = torch.from_numpy(seg).cuda()
    td = torch.from_numpy(direct).cuda().requires_grad_()
    tw = torch.from_numpy(conf).cuda().requires_grad_()
    out = layer([ts, td, tw])
    (out * torch.from_numpy(g).cuda()).sum().backward()
    ro, rd, rw = G.ls_vote_with_grads(seg, direct, conf, g)
    assert np.abs(out.detach().cpu().numpy() - ro).max() < 1e-3
    gd = td.grad.cpu().numpy()
    assert np.isfinite(gd).all() and not gd[0, ys[3], xs[3], 0:2].any()
    _check(gd, rd, "grad_direct")
    _check(tw.grad.cpu().numpy(), rw, "grad_conf")
    gd2, gw2 = layer.backward([ts, td, tw], torch.from_numpy(g).cuda())
    assert torch.equal(gd2, td.grad) and torch.equal(gw2, tw.grad)  # deterministic
