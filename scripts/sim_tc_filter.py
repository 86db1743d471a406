"""CPU end-to-end simulation of the TENSOR-PATH filter planned in DESIGN.md section 7: does "sign of the 3xTF32 value
t, unless |t| is inside the band kappa2 |p| + E, in which case the reference's float32 sequence decides" reproduce the
reference's vote counts bit for bit?  Checked against the golden vote counts of tests/golden/ransac_*.npz (outputs
of the reference's own source, oracle/make_golden.py) for two models of the tensor core's unspecified accumulation
('ideal': exact sum rounded once; 'trunc': every product and partial sum truncated to float32) and for the
evaluation-error bound e1 the plan proposes.  Development tool, no GPU.
usage: python scripts/sim_tc_filter.py [case ...]      (imports oracle/ and tests/golden: test infrastructure)"""
import ast
import math
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle import golden_inputs as GI  # noqa: E402
from oracle import philox_np  # noqa: E402
from oracle import ransac_voting_np as O  # noqa: E402

F = np.float32
u = 2.0 ** -24
E1_TC = 7.9e-6  # 2 sqrt2 x (worst observed truncating error 11.7 u) x 4


def consts(thr):
    th0 = math.acos(thr)
    dC = 1.05 * (u + 8 * u * thr)
    delta = 1.2 * (dC / math.sin(th0) + 8 * u)
    k_lo, k_hi = math.tan(th0 - delta), math.tan(th0 + delta)
    k_lo = float(np.nextafter(np.nextafter(F(k_lo), F(0)), F(0)))
    rho = F(k_hi / k_lo)
    for _ in range(3):
        rho = np.nextafter(rho, F(2))
    kap = F(1.0 - 1.0 / float(rho))
    for _ in range(2):
        kap = np.nextafter(kap, F(1))
    return F(k_lo), float(kap) * 1.001


def tf32(x):
    b = np.asarray(x, F).view(np.uint32).astype(np.uint64)
    return ((b + np.uint64(0x1000)) & np.uint64(0xFFFFE000)).astype(np.uint32).view(F)


def split(x):
    hi = tf32(x)
    return hi.astype(np.float64), tf32((np.asarray(x, F) - hi).astype(F)).astype(np.float64)


def trunc32(x):
    f = np.asarray(x, np.float64).astype(F)
    over = np.abs(f.astype(np.float64)) > np.abs(x)
    return np.where(over, np.nextafter(f, F(0)), f).astype(np.float64)


def form(terms, model):
    if model == "ideal":
        return np.asarray(sum(terms), np.float64).astype(F)
    acc = np.zeros_like(terms[0])
    for t in terms:
        acc = trunc32(acc + trunc32(t))
    return acc.astype(F)


def octn(x, y):
    ax, ay = np.abs(x), np.abs(y)
    return (np.maximum(ax, ay) + F(0.4142136) * np.minimum(ax, ay)) * F(1.0000005)


def exact_votes(direct_v, coords, hyp, thr):
    """reference sequence (voting_for_hypothesis) for hypotheses hyp [n,2] over all pixels -> bool [n, tn]"""
    return O.voting_for_hypothesis(direct_v[:, None, :], coords, hyp[:, None, :], thr)[:, :, 0].astype(bool)


def simulate_job(coords, direct, hyp, thr, model, stats):
    """counts [hn, vn] of the tensor-path filter for one (image, class, round)."""
    k_lo, kappa2 = consts(thr)
    tn, vn, _ = direct.shape
    hn = hyp.shape[0]
    counts = np.zeros((hn, vn), np.int32)
    for v in range(vn):
        dv = direct[:, v]
        h = hyp[:, v]
        q = (dv[:, 0] * dv[:, 0]).astype(F) + (dv[:, 1] * dv[:, 1]).astype(F)
        valid_px = q >= np.uint32(0x2b8cbcce).view(F)
        weird_px = ~((np.abs(dv[:, 0]) <= 2.0 ** 30) & (np.abs(dv[:, 1]) <= 2.0 ** 30))
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = np.where(valid_px, F(1) / np.sqrt(q, dtype=F), F(0)).astype(F)
        D, E = (dv[:, 0] * inv).astype(F), (dv[:, 1] * inv).astype(F)
        G, H = (k_lo * D).astype(F), (k_lo * E).astype(F)
        # hypothesis classes (predicate.cuh: classify_hypothesis)
        fin = np.isfinite(h).all(1)
        zero = ~fin | ~(np.abs((h[:, 0] + h[:, 1]).astype(F)) > F(1e-6))
        big = (np.abs(h) > 2.0 ** 60).any(1)
        with np.errstate(invalid="ignore"):
            near = (np.abs(h - (np.floor(h) + F(0.5))) <= F(8e-6)).all(1) & (np.abs(h) < 4194304.0).all(1)
        exact_h = ~zero & (big | near)
        filt_h = ~zero & ~exact_h
        ex_all = exact_votes(dv, coords, h[exact_h], thr) if exact_h.any() else np.zeros((0, tn), bool)
        counts[exact_h, v] = ex_all.sum(1)
        stats["exact_units"] += int(exact_h.sum()) * tn
        idx_f = np.nonzero(filt_h)[0]
        hf = h[idx_f]
        for c0 in range(0, tn, 128):
            sl = slice(c0, min(tn, c0 + 128))
            cx, cy = coords[sl, 0], coords[sl, 1]
            if weird_px[sl].any():  # whole chunk with the exact predicate
                counts[idx_f, v] += exact_votes(dv[sl], coords[sl], hf, thr).sum(1)
                stats["exact_units"] += len(idx_f) * (sl.stop - sl.start)
                continue
            ox = F(0.5) * F(cx.min() - F(0.5) + cx.max() - F(0.5)) + F(0.5)
            oy = F(0.5) * F(cy.min() - F(0.5) + cy.max() - F(0.5)) + F(0.5)
            cxl, cyl = (cx - ox).astype(F), (cy - oy).astype(F)
            rr = octn(cxl, cyl).max()
            ok = valid_px[sl]
            P0 = ((D[sl] * cyl).astype(F) - (E[sl] * cxl).astype(F)).astype(F)
            A0 = ((G[sl] * cxl).astype(F) + (H[sl] * cyl).astype(F)).astype(F)
            hx, hy = (hf[:, 0] - ox).astype(F), (hf[:, 1] - oy).astype(F)
            Dh, Dl = split(D[sl]); Eh, El = split(E[sl]); Gh, Gl = split(G[sl]); Hh, Hl = split(H[sl])
            Ph, Pl = split(P0); Ah, Al = split(A0)
            xh, xl = split(hx); yh, yl = split(hy)
            xh, xl, yh, yl = xh[:, None], xl[:, None], yh[:, None], yl[:, None]
            z = 0 * xh
            p = form([-Eh[None] * xh, -Eh[None] * xl, -El[None] * xh, Dh[None] * yh, Dh[None] * yl, Dl[None] * yh,
                      -Ph[None] + z, -Pl[None] + z], model)
            s = form([-Gh[None] * xh, -Gh[None] * xl, -Gl[None] * xh, -Hh[None] * yh, -Hh[None] * yl, -Hl[None] * yh,
                      Ah[None] + z, Al[None] + z], model)
            t = (np.abs(p) + s).astype(F)
            Eb = (F(E1_TC) * (octn(hx, hy) + rr)).astype(F)[:, None]
            inband = (np.abs(t) < (F(kappa2) * np.abs(p) + Eb)) & ok[None]
            sign = np.signbit(t) & ok[None]
            if inband.any():
                rows = np.nonzero(inband.any(1))[0]
                ex = exact_votes(dv[sl], coords[sl], hf[rows], thr)
                sub = inband[rows]
                sign[rows] = np.where(sub, ex, sign[rows])
                stats["exact_units"] += int(sub.sum())
            stats["units"] += sign.size
            counts[idx_f, v] += sign.sum(1).astype(np.int32)
    return counts


def run_case(name, models=("ideal", "trunc")):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    if "gen" in g.files:
        mask, vertex = GI.ransac_inputs(**ast.literal_eval(str(g["gen"])))
    else:
        vertex = g["vertex"]
        mask = GI.mask_from_labels(g["labels"], g["points"].shape[1])
    kw = ast.literal_eval(str(g["params"]))
    hn, seed, thr = int(g["hn"]), int(g["seed"]), F(kw.get("inlier_thresh", 0.99))
    max_num = F(kw.get("max_num", 30000))
    b, h, w, oc = mask.shape
    vn = vertex.shape[3]
    for model in models:
        stats = {"units": 0, "exact_units": 0}
        bad = jobs = 0
        for i in range(b):
            for c in range(oc):
                rounds = int(g["rounds"][i, c])
                if rounds == 0:
                    continue
                m = mask[i, :, :, c]
                fg = F(m.sum(dtype=np.float64))
                if fg > max_num:
                    m = m * (philox_np.draw_selection(seed, i, c, h, w) < (max_num / fg)).astype(F)
                ys, xs = np.nonzero(m)
                coords = np.stack([xs, ys], 1).astype(F) + F(0.5)
                direct = vertex[i, ys, xs][:, :, ::-1].astype(F)
                for r in range(rounds):
                    idx = philox_np.draw_idxs(seed, i, c, r, hn, vn, len(ys))
                    hyp = O.generate_hypothesis(direct, coords, idx)
                    cnt = simulate_job(coords, direct, hyp, thr, model, stats)
                    jobs += 1
                    if not np.array_equal(cnt, g["counts_%d_%d_%d" % (i, c, r)]):
                        bad += 1
        print("%-26s %-5s jobs %3d  count mismatches %d  exact fraction %.2e" %
              (name, model, jobs, bad, stats["exact_units"] / max(stats["units"], 1)), flush=True)


if __name__ == "__main__":
    for case in sys.argv[1:] or ["ransac_easy", "ransac_hard", "ransac_cap", "ransac_degenerate"]:
        run_case(case)
