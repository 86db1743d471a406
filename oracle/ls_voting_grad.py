"""Gradient oracle for CoordLSVotingWeighted: the differentiable part of the layer (`calc`,
/root/reference/casapose/pose_estimation/voting_layers_2d.py:83-122, plus the weight activation :32-35)
restated in torch float64 on the CPU and differentiated by torch.autograd — the role TensorFlow's autodiff
plays in the reference's training step (train_casapose.py:536-595).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pinned by tests/golden/ls_grad.npz: torch.autograd over the reference's
own voting_layers_2d.py executed on oracle/tf_standin_torch (TensorFlow's autodiff itself cannot run here).  `hot` (the stop-gradient class weights,
:37-79) comes from oracle.ls_voting_np and is a constant here, as in the reference."""
import numpy as np
import torch

from . import ls_voting_np as OL


def ls_vote_with_grads(seg, direct, w, grad_out, num_points=9, sigmoid_weights=False, filter_estimates=False,
                       output_second_largest_component=False):
    """-> (out [b,oc,vn,2] float64, grad_direct [b,h,w,2vn], grad_w [b,h,w,vn]) for L = sum(out * grad_out)."""
    _, dbg = OL.coord_ls_voting_weighted(seg, direct, w, num_points=num_points, sigmoid_weights=sigmoid_weights,
                                         filter_estimates=filter_estimates,
                                         output_second_largest_component=output_second_largest_component,
                                         return_debug=True)
    hot = torch.from_numpy(dbg["hot"].astype(np.float64))  # [b,h,w,oc]
    b, h, wd, oc = hot.shape
    d = torch.tensor(np.asarray(direct, np.float64).reshape(b, h, wd, num_points, 2), requires_grad=True)
    wl = torch.tensor(np.asarray(w, np.float64), requires_grad=True)
    wgt = torch.sigmoid(wl) if sigmoid_weights else torch.nn.functional.softplus(wl)  # :32-35
    sq = (d * d).sum(-1, keepdim=True)
    norm = torch.sqrt(torch.where(sq > 0, sq, torch.ones_like(sq)))  # :89
    # divide_no_nan :90.  An exactly zero vector gets a ZERO gradient here; TensorFlow's sqrt gradient would give
    # 0 * inf = NaN for it (network outputs are never exactly zero, synthetic backgrounds are).
    n = torch.where(sq > 0, d / norm, torch.zeros_like(d))
    eye = torch.eye(2, dtype=torch.float64)
    R = (eye - n[..., :, None] * n[..., None, :]) * wgt[..., None, None]  # [b,h,w,vn,2,2]  :92-94
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float64), torch.arange(wd, dtype=torch.float64), indexing="ij")
    p = torch.stack([(ys + 0.5) / h, (xs + 0.5) / h], -1)[None, :, :, None, :]  # :95-99 both over the height
    q = R[..., 0] * p[..., 0:1] + R[..., 1] * p[..., 1:2]  # :103-105
    Rc = torch.einsum("bhwkij,bhwc->bckij", R, hot)  # :108, :113
    qc = torch.einsum("bhwki,bhwc->bcki", q, hot)  # :107, :114
    rcond = 10.0 * 2 * np.finfo(np.float64).eps
    pinv = torch.linalg.pinv(Rc, rtol=rcond)  # :116
    out = (pinv @ qc[..., None])[..., 0] * h  # :120-122
    loss = (out * torch.from_numpy(np.asarray(grad_out, np.float64))).sum()
    loss.backward()
    return out.detach().numpy(), d.grad.reshape(b, h, wd, 2 * num_points).numpy(), wl.grad.numpy()
