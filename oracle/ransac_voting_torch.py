"""torch-CPU twin of the numpy oracle — the timed "CPU restatement of the reference".

TEST / BENCH INFRASTRUCTURE (see oracle/__init__.py); bit-identical to the numpy oracle, which tests/golden/ pins.

Written independently of oracle/ransac_voting_np.py, one torch op per TensorFlow op of
/root/reference/casapose/pose_estimation/ransac_voting.py:197-368, and — like the reference —
it MATERIALISES the [hn, tn, vn] temporaries of voting_for_hypothesis (:232-247) instead of
fusing them.  It is what `bench.py --impl reference` and the `cpu_baseline` leg time on the
host cores (torch intra-op threads = all cores), and tests/test_oracle_twin.py checks that its
vote counts are bit-identical to the numpy oracle's.  Label in every report:
"CPU restatement of the reference", never "reference TF".
"""
import torch

from . import philox_np
from .ransac_voting_np import is_invertible, stop_test

EPS = 1e-6


def generate_hypothesis(direct, coords, idxs):
    """:197-227"""
    hn, vn, _ = idxs.shape
    v_idx = torch.arange(vn).expand(hn, vn)
    c_s = coords[idxs]  # tf.gather                                  :216
    d_s = torch.stack([direct[idxs[:, :, 0], v_idx], direct[idxs[:, :, 1], v_idx]], dim=2)  # tf.gather_nd  :217
    det = d_s[:, :, 1, 0] * d_s[:, :, 0, 1] - d_s[:, :, 1, 1] * d_s[:, :, 0, 0]
    u = ((c_s[:, :, 1, 1] - c_s[:, :, 0, 1]) * d_s[:, :, 1, 0] - (c_s[:, :, 1, 0] - c_s[:, :, 0, 0]) * d_s[:, :, 1, 1]) / det
    hypo = c_s[:, :, 0] + d_s[:, :, 0] * u.unsqueeze(2)
    return torch.where((det.abs() > EPS).unsqueeze(2), hypo, torch.zeros((), dtype=hypo.dtype))


def voting_for_hypothesis(direct, coords, cur_hyp_pts, inlier_thresh):
    """:230-249 — full [hn, tn, vn(,2)] temporaries, as the reference's graph builds them."""
    coords = coords.unsqueeze(1).unsqueeze(0)  # [1,tn,1,2]
    direct = direct.unsqueeze(0)  # [1,tn,vn,2]
    cur_hyp_pts = cur_hyp_pts.unsqueeze(1)  # [hn,1,vn,2]
    hypo_dirs = cur_hyp_pts - coords  # [hn,tn,vn,2]
    norm_dir = (direct * direct).sum(-1).sqrt()
    norm_hyp = (hypo_dirs * hypo_dirs).sum(-1).sqrt()
    valid = (norm_dir > EPS) & (norm_hyp > EPS)
    valid = valid & (cur_hyp_pts.sum(-1).abs() > EPS)
    angle = (direct * hypo_dirs).sum(-1) / (norm_dir * norm_hyp)
    return torch.where(valid & (angle > inlier_thresh), 1, 0).to(torch.int32)


def ransac_voting_batch(cur_mask, cur_vertex, inlier_thresh, confidence, max_iter, min_num, max_num, round_hyp_num, vn,
                        seed=0, image=0, cls=0, max_rounds=None):
    """:275-368 for one (image, class).  Returns (points [vn,2], info dict)."""
    h, w = cur_mask.shape
    info = {"tn": 0, "rounds": 0, "counts": [], "units": 0}
    foreground_num = cur_mask.sum()
    if foreground_num < min_num:
        return torch.zeros(vn, 2), info
    if foreground_num > max_num:
        selection = torch.from_numpy(philox_np.draw_selection(seed, image, cls, h, w))
        cur_mask = cur_mask * (selection < (max_num / foreground_num)).float()
    coords = torch.nonzero(cur_mask != 0.0).flip(1).float() + 0.5
    direct = cur_vertex[cur_mask.bool()].flip(2)
    tn = coords.shape[0]
    info["tn"] = tn
    if tn == 0:
        return torch.zeros(vn, 2), info
    all_win_ratio = torch.zeros(vn)
    all_win_pts = torch.zeros(vn, 2)
    cur_iter, hyp_num = 0, 0
    while True:
        idxs = torch.from_numpy(philox_np.draw_idxs(seed, image, cls, cur_iter, round_hyp_num, vn, tn)).long()
        cur_hyp_pts = generate_hypothesis(direct, coords, idxs)
        cur_inlier = voting_for_hypothesis(direct, coords, cur_hyp_pts, inlier_thresh)
        counts = cur_inlier.sum(1, dtype=torch.int32)
        cur_win_idx = torch.argmax(counts, 0)
        cur_win_counts = counts.max(0).values
        cur_win_pts = cur_hyp_pts[cur_win_idx, torch.arange(vn)]
        cur_win_ratio = cur_win_counts.float() / float(tn)
        larger = all_win_ratio < cur_win_ratio
        all_win_pts = torch.where(larger.unsqueeze(1), cur_win_pts, all_win_pts)
        all_win_ratio = torch.where(larger, cur_win_ratio, all_win_ratio)
        hyp_num += round_hyp_num
        cur_iter += 1
        info["counts"].append(counts)
        info["units"] += round_hyp_num * tn * vn
        if stop_test(all_win_ratio.min().item(), hyp_num, confidence) or cur_iter >= max_iter:
            break
        if max_rounds is not None and cur_iter >= max_rounds:
            break
    info["rounds"] = cur_iter
    normal = (direct * torch.tensor([1.0, -1.0])).flip(2)
    all_inlier = voting_for_hypothesis(direct, coords, all_win_pts.unsqueeze(0), inlier_thresh)[0].float()
    info["units"] += tn * vn
    normal = (normal * all_inlier.unsqueeze(2)).permute(1, 0, 2)  # [vn,tn,2]
    b = (normal * coords.unsqueeze(0)).sum(2)
    ata = torch.matmul(normal.permute(0, 2, 1).double(), normal.double()).float()  # float64 accumulation (see numpy oracle)
    atb = (normal * b.unsqueeze(2)).double().sum(1).float()
    ok = all(is_invertible(float(ata[v, 0, 0]), float(ata[v, 0, 1]), float(ata[v, 1, 1])) for v in range(vn))
    if not ok:
        return all_win_pts, info
    a, bb, c = ata[:, 0, 0].double(), ata[:, 0, 1].double(), ata[:, 1, 1].double()
    g0, g1 = atb[:, 0].double(), atb[:, 1].double()
    det = a * c - bb * bb
    pts = torch.stack([(c * g0 - bb * g1) / det, (a * g1 - bb * g0) / det], dim=1).float()
    return pts, info


def ransac_voting_layer_all_masks(mask, vertex, round_hyp_num, inlier_thresh=0.99, confidence=0.99, max_iter=20,
                                  min_num=5, max_num=30000, seed=0, image_offset=0, return_info=False):
    """:446-484.  mask [b,h,w,oc], vertex [b,h,w,vn,2] CPU float32 tensors -> [b,oc,vn,2]."""
    b, h, w, oc = mask.shape
    vn = vertex.shape[3]
    out = torch.zeros(b, oc, vn, 2)
    infos = []
    for i in range(b):
        m = mask[i].permute(2, 0, 1)  # :429
        for c in range(oc):
            pts, info = ransac_voting_batch(m[c], vertex[i], inlier_thresh, confidence, max_iter, float(min_num),
                                            float(max_num), int(round_hyp_num), vn, seed=seed, image=image_offset + i, cls=c)
            out[i, c] = pts
            infos.append(info)
    return (out, infos) if return_info else out
