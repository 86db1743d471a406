"""numpy restatement of the reference's weighted least-squares keypoint layer.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pinned by tests/golden/ls_*.npz — the reference's own
voting_layers_2d.py executed over a numpy TensorFlow stand-in (oracle/make_golden.py); TensorFlow and
tensorflow-addons themselves are not installable here.

Follows /root/reference/casapose/pose_estimation/voting_layers_2d.py:5-122 op by op in float32,
with the float64 accumulation the reference itself prescribes (:113-114).  Third-party pieces:
  * tfa.image.connected_components (tensorflow-addons 0.17.0, :53) — source not vendored in the
    reference; restated from its published contract (4-connectivity, ids 1..n in row-major order
    of each component's first pixel, 0 for zero pixels) with scipy.ndimage.label, whose default
    structure and id order are the same;
  * tf.linalg.pinv (:116) — numpy.linalg.pinv with TensorFlow's default rcond
    10 * max(rows, cols) * eps(float64);
  * tf.math.top_k tie rule (:67): equal values keep the lower index first — a stable argsort.
Elementwise float32 results may differ from TensorFlow/XLA in the last ulp (exp, log1p, FMA
contraction under jit_compile); the sums are float64, so the layer's output is compared at 1e-3 px.
"""
import numpy as np
from scipy import ndimage

F32 = np.float32


def softplus_f32(x):
    """tf.math.softplus (:35), Eigen's functor: x if x > T, exp(x) if x < -T, else log(exp(x) + 1); T = -(log(eps) + 2)."""
    x = x.astype(F32)
    thr = F32(-(np.log(np.finfo(np.float32).eps) + 2.0))  # 13.942385
    with np.errstate(over="ignore"):
        ex = np.exp(x).astype(F32)
        mid = np.log(ex + F32(1.0)).astype(F32)
    return np.where(x > thr, x, np.where(x < -thr, ex, mid)).astype(F32)


def sigmoid_f32(x):
    with np.errstate(over="ignore"):
        return (F32(1.0) / (F32(1.0) + np.exp(-x.astype(F32)))).astype(F32)  # :33


def hard_softmax_f32(seg):
    """softmax(seg * 1e6) in float32 (:38-41)."""
    z = seg.astype(F32) * F32(1e6)
    z = z - z.max(axis=-1, keepdims=True)
    with np.errstate(under="ignore"):
        e = np.exp(z).astype(F32)
    return (e / e.sum(axis=-1, keepdims=True, dtype=F32)).astype(F32)


def select_component(hot_int, bins, which):
    """:51-76 for one [h,w] int image: returns the float 0/1 map `components == top_k index[which]`."""
    comp, _ = ndimage.label(hot_int)  # 4-connectivity, raster-order ids
    comp = comp.astype(np.int32).reshape(-1)
    bincount = np.bincount(comp, minlength=bins)  # :64
    bincount = np.where(bincount < 50, 0, bincount)  # :66
    order = np.argsort(-bincount, kind="stable")  # :67 top_k: descending, ties -> lower index first
    return (comp == order[which]).astype(F32)


def coord_ls_voting_weighted(
    seg,
    direct,
    w,
    num_classes=None,
    num_points=9,
    sigmoid_weights=False,
    filter_estimates=False,
    output_second_largest_component=False,
    return_debug=False,
):
    """CoordLSVotingWeighted(...)([seg, direct, w]) -> [b, oc, num_points, 2] (y, x) pixels, float32.

    seg [b,h,w,1+oc] logits, direct [b,h,w,2*num_points] (dy,dx per keypoint), w [b,h,w,num_points] logits.
    """
    seg = np.asarray(seg, F32)
    direct = np.asarray(direct, F32)
    w = np.asarray(w, F32)
    b, h, wd, nc = seg.shape
    oc = nc - 1
    wgt = sigmoid_f32(w) if sigmoid_weights else softplus_f32(w)  # :32-35
    hot = hard_softmax_f32(seg)[..., 1:]  # [b,h,w,oc]                         :39-41
    if filter_estimates:  # :43-79
        hot_int = (hot + F32(0.1)).astype(np.int32)  # :44
        bins, which = (3, 2) if output_second_largest_component else (2, 1)
        keep = np.zeros_like(hot)
        for i in range(b):
            for c in range(oc):
                keep[i, :, :, c] = select_component(hot_int[i, :, :, c], bins, which).reshape(h, wd)
        hot = keep * hot  # :79

    # calc (:83-122)
    n = direct.reshape(b, h, wd, num_points, 2)
    norm = np.sqrt(n[..., 0] * n[..., 0] + n[..., 1] * n[..., 1])[..., None]  # :89
    with np.errstate(divide="ignore", invalid="ignore"):
        n = np.where(norm != 0, n / norm, F32(0.0)).astype(F32)  # divide_no_nan :90
    n0, n1 = n[..., 0], n[..., 1]
    R00 = (F32(1.0) - n0 * n0) * wgt  # :92-94
    R01 = (F32(0.0) - n0 * n1) * wgt
    R10 = (F32(0.0) - n1 * n0) * wgt
    R11 = (F32(1.0) - n1 * n1) * wgt
    ys, xs = np.meshgrid(np.arange(h), np.arange(wd), indexing="ij")
    cy = ((ys.astype(F32) + F32(0.5)) / F32(h))[None, :, :, None]  # :97  both axes divided by the height
    cx = ((xs.astype(F32) + F32(0.5)) / F32(h))[None, :, :, None]  # :96
    q0 = R00 * cy + R01 * cx  # :103-105
    q1 = R10 * cy + R11 * cx

    out = np.zeros((b, oc, num_points, 2), F32)
    Rc = np.zeros((b, oc, num_points, 2, 2), np.float64)
    qc = np.zeros((b, oc, num_points, 2), np.float64)
    rcond = 10.0 * 2 * np.finfo(np.float64).eps
    for c in range(oc):
        hc = hot[..., c][..., None]  # [b,h,w,1]
        sel = hc != 0  # multiply_no_nan: y == 0 -> 0 even for non-finite x (:107-108)
        for name, src in (("00", R00), ("01", R01), ("10", R10), ("11", R11)):
            val = np.where(sel, src * hc, F32(0.0)).astype(np.float64).sum(axis=(1, 2))  # :113
            Rc[:, c, :, int(name[0]), int(name[1])] = val
        qc[:, c, :, 0] = np.where(sel, q0 * hc, F32(0.0)).astype(np.float64).sum(axis=(1, 2))  # :114
        qc[:, c, :, 1] = np.where(sel, q1 * hc, F32(0.0)).astype(np.float64).sum(axis=(1, 2))
    for i in range(b):
        for c in range(oc):
            for k in range(num_points):
                pinv = np.linalg.pinv(Rc[i, c, k], rcond=rcond)  # :116
                p = pinv @ qc[i, c, k]  # :120
                out[i, c, k] = p.astype(F32) * F32(h)  # :122
    if return_debug:
        return out, {"hot": hot, "R": Rc, "q": qc, "weights": wgt}
    return out
