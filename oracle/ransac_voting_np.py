"""numpy restatement of the reference's RANSAC keypoint voting.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pinned by tests/golden/ (the reference's own source run over a
numpy TensorFlow stand-in, oracle/make_golden.py); TensorFlow's kernels themselves cannot run here.

Follows /root/reference/casapose/pose_estimation/ransac_voting.py op by op:
every TensorFlow op is one numpy float32 op, so every intermediate is rounded to
float32 exactly once, in the order the reference's graph evaluates it (separate TF
kernels never fuse a multiply into an add).

Choices where TensorFlow's behaviour is implementation-defined (documented in
DESIGN.md, all shared verbatim with the CUDA path):
  * random numbers: explicit Philox streams (oracle/philox_np.py) replace
    tf.random.uniform (:296, :319) — or caller-supplied ``idxs`` / ``selection``.
  * ``x ** hyp_num`` in the stop test (:345): binary exponentiation in float64 on the
    float32 base, rounded once to float32 (``_pow_f32``); identical instruction
    sequence on host and device, within 1 ulp of Eigen's powf.
  * sums over the ``tn`` pixel axis in the refinement (:361-362): float32 products,
    float64 accumulation, one final rounding to float32 (TF's float32 summation
    order is a property of Eigen's GEMM/reduction kernels, not of the algorithm).
  * 2x2 condition number and inverse (:254-272, :367): closed form in float64 on the
    float32 ATA / ATb (TF uses LAPACK-style SVD / LU in float32).
"""
import math

import numpy as np

from . import philox_np

F32 = np.float32
EPS_1E6 = F32(1e-6)


def generate_hypothesis(direct, coords, idxs):
    """ransac_voting.py:197-227.  direct [tn,vn,2] (dx,dy), coords [tn,2] (x,y), idxs [hn,vn,2] -> [hn,vn,2]."""
    hn, vn, _ = idxs.shape
    v_idx = np.broadcast_to(np.arange(vn), (hn, vn))
    c_s = coords[idxs]  # [hn,vn,2(pair),2(xy)]                                   :216
    d_s = np.stack([direct[idxs[:, :, 0], v_idx], direct[idxs[:, :, 1], v_idx]], axis=2)  # :217

    det = d_s[:, :, 1, 0] * d_s[:, :, 0, 1] - d_s[:, :, 1, 1] * d_s[:, :, 0, 0]  # :219
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        u = (
            (c_s[:, :, 1, 1] - c_s[:, :, 0, 1]) * d_s[:, :, 1, 0]
            - (c_s[:, :, 1, 0] - c_s[:, :, 0, 0]) * d_s[:, :, 1, 1]
        ) / det  # :221-223
        hypo_pts = c_s[:, :, 0] + d_s[:, :, 0] * u[:, :, None]  # :225
    hypo_pts = np.where((np.abs(det) > EPS_1E6)[:, :, None], hypo_pts, F32(0.0))  # :226
    return hypo_pts.astype(F32)


def voting_for_hypothesis(direct, coords, cur_hyp_pts, inlier_thresh, chunk=32):
    """ransac_voting.py:230-249.  Returns int32 inlier flags [hn,tn,vn].

    Evaluated in chunks over hn only to bound memory; the arithmetic is elementwise,
    so chunking cannot change a single bit.
    """
    thr = F32(inlier_thresh)
    hn = cur_hyp_pts.shape[0]
    tn, vn, _ = direct.shape
    out = np.empty((hn, tn, vn), dtype=np.int32)
    co = coords[None, :, None, :]  # [1,tn,1,2]      :232
    di = direct[None]  # [1,tn,vn,2]                  :233
    norm_dir = np.sqrt(di[..., 0] * di[..., 0] + di[..., 1] * di[..., 1])  # :238  tf.norm = sqrt(sum(x*x))
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        for s in range(0, hn, chunk):
            hp = cur_hyp_pts[s : s + chunk, None]  # [c,1,vn,2]              :235
            hd = hp - co  # [c,tn,vn,2]                                      :236
            norm_hyp = np.sqrt(hd[..., 0] * hd[..., 0] + hd[..., 1] * hd[..., 1])  # :239
            valid = (norm_dir > EPS_1E6) & (norm_hyp > EPS_1E6)  # :240
            valid = valid & (np.abs(hp[..., 0] + hp[..., 1]) > EPS_1E6)  # :241-243
            dot = di[..., 0] * hd[..., 0] + di[..., 1] * hd[..., 1]
            ang = dot / (norm_dir * norm_hyp)  # :245
            out[s : s + chunk] = (valid & (ang > thr)).astype(np.int32)  # :247
    return out


def _ipow_f64(x, n):
    """x**n for integer n >= 0 by binary exponentiation; pure float64 multiplies (no FMA)."""
    result = 1.0
    b = float(x)
    n = int(n)
    while n:
        if n & 1:
            result = result * b
        b = b * b
        n >>= 1
    return result


def _pow_f32(base_f32, n):
    with np.errstate(under="ignore"):
        return F32(_ipow_f64(float(F32(base_f32)), n))


def stop_test(cur_min_ratio, hyp_num, confidence):
    """ransac_voting.py:344-346:  1 - (1 - r**2) ** hyp_num > confidence  (all float32)."""
    r = F32(cur_min_ratio)
    r2 = F32(r * r)  # tf.pow(r, 2.0) is the correctly rounded square
    base = F32(F32(1.0) - r2)
    pw = _pow_f32(base, hyp_num)
    return bool(F32(F32(1.0) - pw) > F32(confidence))


def cond_2x2_sym_f64(a, b, c):
    """Condition number s0/s1 of [[a,b],[b,c]] (ransac_voting.py:254-264), closed form in float64."""
    a, b, c = float(a), float(b), float(c)
    m = (a + c) * 0.5
    d = (a - c) * 0.5
    r = math.sqrt(d * d + b * b)
    s0 = abs(m + r)
    s1 = abs(m - r)
    if s1 > s0:
        s0, s1 = s1, s0
    if s1 == 0.0:
        return math.inf if s0 != 0.0 else math.nan
    return s0 / s1


def is_invertible(a, b, c, epsilon=1e-6):
    """ransac_voting.py:267-272."""
    eps_inv = float(F32(1.0 / epsilon))
    cnd = cond_2x2_sym_f64(a, b, c)
    return math.isfinite(cnd) and cnd < eps_inv


def ransac_voting_batch(
    cur_mask,
    cur_vertex,
    inlier_thresh,
    confidence,
    max_iter,
    min_num,
    max_num,
    round_hyp_num,
    vn,
    *,
    seed=0,
    image=0,
    cls=0,
    idxs=None,
    selection=None,
    accumulate="float64",
):
    """ransac_voting.py:275-368 for one (image, class).

    cur_mask [h,w] float32, cur_vertex [h,w,vn,2] float32 (dy,dx).
    idxs: optional int32 [rounds,hn,vn,2] (used instead of the Philox stream),
    selection: optional float32 [h,w].
    Returns a dict: points [vn,2] (x,y) plus every intermediate the parity tests compare.
    """
    cur_mask = np.asarray(cur_mask, dtype=F32)
    cur_vertex = np.asarray(cur_vertex, dtype=F32)
    h, w = cur_mask.shape
    hn = int(round_hyp_num)
    res = {
        "points": np.zeros((vn, 2), F32),
        "win_pts": np.zeros((vn, 2), F32),
        "tn0": 0,
        "tn": 0,
        "rounds": 0,
        "counts": [],
        "win_idx": [],
        "hyps": [],
        "refined": False,
    }
    # :287 tf.reduce_sum(cur_mask).  For the {0,1} masks the callers build (one_hot, pose_evaluation.py:37)
    # this is the exact pixel count (< 2**24); summed in float64 so the oracle does not depend on numpy's order.
    foreground_num = F32(cur_mask.sum(dtype=np.float64))
    res["tn0"] = int(np.count_nonzero(cur_mask))
    if foreground_num < F32(min_num):  # :290-292
        return res
    if foreground_num > F32(max_num):  # :295-301
        if selection is None:
            selection = philox_np.draw_selection(seed, image, cls, h, w)
        keep = selection.astype(F32) < (F32(max_num) / foreground_num)
        cur_mask = cur_mask * keep.astype(F32)

    ys, xs = np.nonzero(cur_mask != 0.0)  # raster order, like tf.where      :303-305
    coords = np.stack([xs, ys], axis=1).astype(F32) + F32(0.5)  # (x,y)+0.5      :306
    direct = cur_vertex[ys, xs][:, :, ::-1].astype(F32)  # [tn,vn,2] (dx,dy)     :308
    tn = coords.shape[0]
    res["tn"] = tn
    if tn == 0:
        return res

    all_win_ratio = np.zeros(vn, F32)
    all_win_pts = np.zeros((vn, 2), F32)
    cur_iter = 0
    hyp_num = F32(0.0)
    while True:  # :318
        if idxs is not None:
            cur_idx = np.asarray(idxs[cur_iter], dtype=np.int32)
        else:
            cur_idx = philox_np.draw_idxs(seed, image, cls, cur_iter, hn, vn, tn)
        cur_hyp_pts = generate_hypothesis(direct, coords, cur_idx)  # :322
        cur_inlier = voting_for_hypothesis(direct, coords, cur_hyp_pts, inlier_thresh)  # :324
        cur_inlier_counts = cur_inlier.sum(axis=1, dtype=np.int32)  # [hn,vn]  :327
        cur_win_idx = np.argmax(cur_inlier_counts, axis=0).astype(np.int32)  # first max  :328
        cur_win_counts = cur_inlier_counts.max(axis=0)  # :330
        cur_win_pts = cur_hyp_pts[cur_win_idx, np.arange(vn)]  # :332
        cur_win_ratio = cur_win_counts.astype(F32) / F32(tn)  # :333
        larger = all_win_ratio < cur_win_ratio  # :336
        all_win_pts = np.where(larger[:, None], cur_win_pts, all_win_pts)  # :337
        all_win_ratio = np.where(larger, cur_win_ratio, all_win_ratio)  # :338
        hyp_num = F32(hyp_num + F32(hn))  # :340
        cur_iter += 1  # :341
        res["counts"].append(cur_inlier_counts)
        res["win_idx"].append(cur_win_idx)
        res["hyps"].append(cur_hyp_pts)
        cur_min_ratio = all_win_ratio.min()  # :342
        if stop_test(cur_min_ratio, int(hyp_num), confidence) or cur_iter >= int(max_iter):  # :344-347
            break
    res["rounds"] = cur_iter
    res["win_pts"] = all_win_pts.astype(F32)
    res["win_ratio"] = all_win_ratio

    normal = (direct * np.array([1, -1], F32))[:, :, ::-1]  # (-dy, dx) [tn,vn,2]  :349
    all_inlier = voting_for_hypothesis(direct, coords, all_win_pts[None], inlier_thresh)[0]  # [tn,vn] :353
    res["inlier"] = all_inlier
    normal = normal * all_inlier.astype(F32)[:, :, None]  # :356
    normal = normal.transpose(1, 0, 2)  # [vn,tn,2]                              :357
    nc = normal * coords[None]  # float32 products
    b = nc[..., 0] + nc[..., 1]  # [vn,tn]                                        :359
    acc = np.float64 if accumulate == "float64" else np.float32
    pxx = normal[..., 0] * normal[..., 0]
    pxy = normal[..., 0] * normal[..., 1]
    pyy = normal[..., 1] * normal[..., 1]
    ata = np.stack([pxx.sum(1, dtype=acc), pxy.sum(1, dtype=acc), pyy.sum(1, dtype=acc)], axis=1).astype(F32)  # :361
    nb = normal * b[..., None]
    atb = nb.sum(1, dtype=acc).astype(F32)  # [vn,2]                              :362
    res["ata"] = ata
    res["atb"] = atb
    inv_ok = np.array([is_invertible(*ata[v]) for v in range(vn)])
    res["invertible"] = inv_ok
    if not inv_ok.all():  # :364-365
        res["points"] = all_win_pts.astype(F32)
        return res
    pts = np.zeros((vn, 2), F32)
    for v in range(vn):  # :367  inv(ATA) @ ATb, closed form in float64
        a, bb, c = (float(t) for t in ata[v])
        g0, g1 = float(atb[v, 0]), float(atb[v, 1])
        det = a * c - bb * bb
        pts[v, 0] = F32((c * g0 - bb * g1) / det)
        pts[v, 1] = F32((a * g1 - bb * g0) / det)
    res["points"] = pts
    res["refined"] = True
    return res


def ransac_voting_layer_all_masks(
    mask,
    vertex,
    round_hyp_num,
    inlier_thresh=0.99,
    confidence=0.99,
    max_iter=20,
    min_num=5,
    max_num=30000,
    *,
    seed=0,
    image_offset=0,
    idxs=None,
    selection=None,
    return_debug=False,
    accumulate="float64",
):
    """ransac_voting.py:446-484 (+ :410-443).  mask [b,h,w,oc], vertex [b,h,w,vn,2] -> [b,oc,vn,2] (x,y).

    idxs: optional int32 [b,oc,rounds,hn,vn,2]; selection: optional float32 [b,oc,h,w].
    """
    mask = np.asarray(mask, dtype=F32)
    vertex = np.asarray(vertex, dtype=F32)
    b, h, w, oc = mask.shape
    vn = vertex.shape[3]
    out = np.zeros((b, oc, vn, 2), F32)
    debug = [[None] * oc for _ in range(b)]
    for i in range(b):  # tf.map_fn over images :483
        for c in range(oc):  # tf.map_fn over classes :442 (mask transposed to [oc,h,w] :429)
            r = ransac_voting_batch(
                mask[i, :, :, c],
                vertex[i],
                inlier_thresh,
                confidence,
                max_iter,
                min_num,
                max_num,
                round_hyp_num,
                vn,
                seed=seed,
                image=image_offset + i,
                cls=c,
                idxs=None if idxs is None else idxs[i, c],
                selection=None if selection is None else selection[i, c],
                accumulate=accumulate,
            )
            out[i, c] = r["points"]
            debug[i][c] = r
    if return_debug:
        return out, debug
    return out
