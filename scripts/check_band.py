"""Executable check of the filtered predicate's error band (casapose_b200/csrc/predicate.cuh: filter_consts()).

No GPU, numpy only.  For every threshold and spread it draws adversarial units concentrated on the decision
boundary (hypothesis at angle theta0 +- eps from the stored direction, eps down to 1e-8 rad, distances 2^-4 .. 2^14 px,
direction magnitudes 2^-10 .. 2^10, chunk origin up to 90 px from the pixel) and evaluates

  ref     the reference's float32 sequence (ransac_voting.py:236-247), one rounding per op      -> verdict, ang
  truth   cos(theta) and t* = |hd| sin(theta - theta0)/cos(theta0) in float64 on the reference's hd = fl(h - c)
  filt    the chunk-local form of k_score in emulated float32 (FMA = one rounding of the float64 a*b+c; rsqrt
          perturbed by up to +-2 ulp; both contraction choices of P0 / A0)

and asserts the three statements the derivation in predicate.cuh rests on:

  (1) |ang - cos(theta)| <= u + 7u|cos(theta)| + 100u^2 <= dC           (the reference's own rounding)
  (4) |t_eval - t*| <= e1 (|h'| + R) + (d_fil / cos(theta0)) |hd|        (evaluation error + coefficient rounding)
  (5) sign(t_eval) == verdict   whenever |t_eval| >= c1 (|h'| + R)       (what the main loop relies on)
      sign(t_eval) == verdict   whenever |t_eval| >= kappa2 |p| + e1f (|h'| + R)   (what stage 2 relies on)
  (6) the direct form of the refinement's re-vote (k_refine_solve: hd = fl(h - c), no chunk origin):
      sign(t_direct) == verdict whenever |t_direct| >= c1 (|hdx| + |hdy|)

with R = this pixel's own offset from the chunk origin (the smallest radius a chunk containing it can have: the
tightest bound the kernel can ever apply).  It prints the observed maxima next to the budgets, i.e. the margins.
usage: python scripts/check_band.py [samples per (threshold, spread), default 400000]
"""
import math
import sys

import numpy as np

F = np.float32
U = 2.0 ** -24


def consts(thr32):
    thr = float(thr32)
    th0 = math.acos(thr)
    s0 = math.sin(th0)
    dC = 1.05 * (U + 8.0 * U * thr)
    d_ref, d_fil = dC / s0, 8.0 * U
    delta = 1.2 * (d_ref + d_fil)
    assert th0 - delta > 1e-3
    assert math.cos(th0 - delta) >= thr + dC and math.cos(th0 + delta) <= thr - dC
    w = math.sin(delta) / math.cos(th0)
    kappa2 = w / math.sin(th0 - delta)
    e1 = 12.0 * 1.41421357 * U
    return dict(thr=thr, th0=th0, dC=dC, delta=delta, d_fil=d_fil, w=w, e1_raw=e1, k_mid=F(math.tan(th0)),
                e1=F(1.001 * (1.0 + kappa2) * e1), kappa2=F(1.001 * kappa2), c1=F(1.001 * (w + e1)))


def fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


def octn(x, y):
    ax, ay = np.abs(x), np.abs(y)
    return (fma32(np.full_like(ax, F(0.4142136)), np.minimum(ax, ay), np.maximum(ax, ay)) * F(1.0000005)).astype(F)


def run(thr32, spread, n, rng):
    k = consts(thr32)
    cx = (rng.integers(0, 1920, n) + 0.5).astype(F)
    cy = (rng.integers(0, 1080, n) + 0.5).astype(F)
    phi = rng.uniform(0, 2 * math.pi, n)
    mag = np.exp2(rng.uniform(-10, 10, n))
    dx, dy = (mag * np.cos(phi)).astype(F), (mag * np.sin(phi)).astype(F)
    sgn = rng.choice([-1.0, 1.0], n)
    th = k["th0"] + spread * rng.uniform(-1, 1, n)
    dist = np.exp2(rng.uniform(-4, 14, n))
    phir = np.arctan2(dy.astype(np.float64), dx.astype(np.float64))
    hx = (cx.astype(np.float64) + dist * np.cos(phir + sgn * th)).astype(F)
    hy = (cy.astype(np.float64) + dist * np.sin(phir + sgn * th)).astype(F)

    # ---- the reference sequence, float32, one rounding per op
    hdx, hdy = (hx - cx).astype(F), (hy - cy).astype(F)
    nd = np.sqrt((dx * dx + dy * dy).astype(F)).astype(F)
    nh = np.sqrt((hdx * hdx + hdy * hdy).astype(F)).astype(F)
    dot = ((dx * hdx).astype(F) + (dy * hdy).astype(F)).astype(F)
    ang = (dot / (nd * nh).astype(F)).astype(F)
    valid = (nd > F(1e-6)) & (nh > F(1e-6)) & (np.abs((hx + hy).astype(F)) > F(1e-6))
    verdict = valid & (ang > F(thr32))

    # ---- truth in float64 on the reference's own hd
    d64 = np.stack([dx, dy], 1).astype(np.float64)
    h64 = np.stack([hdx, hdy], 1).astype(np.float64)
    nd64, nh64 = np.hypot(d64[:, 0], d64[:, 1]), np.hypot(h64[:, 0], h64[:, 1])
    keep = valid & (nh64 > 1e-3)
    cos_t = (d64 * h64).sum(1) / (nd64 * nh64)
    cross = (d64[:, 0] * h64[:, 1] - d64[:, 1] * h64[:, 0]) / nd64
    dotn = (d64 * h64).sum(1) / nd64
    t_star = np.abs(cross) - math.tan(k["th0"]) * dotn
    err1 = np.abs(ang.astype(np.float64) - cos_t)[keep]
    bound1 = (U + 7 * U * np.abs(cos_t) + 100 * U * U)[keep]
    assert (err1 <= bound1).all(), "(1) reference rounding exceeds u + 7u|cos|"
    assert (bound1 <= k["dC"]).all() or spread > 1e-2, "(1) dC does not cover the analytic bound"

    # ---- the chunk-local form of k_score in emulated float32
    ox = cx + F(0.5) * rng.integers(-180, 181, n).astype(F)
    oy = cy + F(0.5) * rng.integers(-180, 181, n).astype(F)
    cxl, cyl = (cx - ox).astype(F), (cy - oy).astype(F)  # exact
    q = (dx * dx + dy * dy).astype(F)
    inv = (1.0 / np.sqrt(q.astype(np.float64))).astype(F)
    steps = rng.integers(-2, 3, n)  # rsqrtf: 2 ulp
    inv = np.where(steps > 0, np.nextafter(inv, F(np.inf)), np.where(steps < 0, np.nextafter(inv, F(0)), inv))
    inv = np.where(np.abs(steps) > 1, np.where(steps > 0, np.nextafter(inv, F(np.inf)), np.nextafter(inv, F(0))), inv).astype(F)
    D, E = (dx * inv).astype(F), (dy * inv).astype(F)
    G, H = (k["k_mid"] * D).astype(F), (k["k_mid"] * E).astype(F)
    contr = rng.integers(0, 2, n).astype(bool)  # nvcc may or may not contract a*b - c*d / a*b + c*d
    P0 = np.where(contr, fma32(D, cyl, -(E * cxl).astype(F)), ((D * cyl).astype(F) - (E * cxl).astype(F)).astype(F)).astype(F)
    A0 = np.where(contr, fma32(G, cxl, (H * cyl).astype(F)), ((G * cxl).astype(F) + (H * cyl).astype(F)).astype(F)).astype(F)
    hxl, hyl = (hx - ox).astype(F), (hy - oy).astype(F)  # h' = fl(h - o)
    p = fma32(D, hyl, fma32(-E, hxl, -P0))
    s = fma32(-G, hxl, fma32(-H, hyl, A0))
    t = (np.abs(p) + s).astype(F)
    nrm = (octn(hxl, hyl) + octn(cxl, cyl)).astype(F)  # |h'| + R, R = this pixel's own offset

    hp_true = np.hypot(hxl.astype(np.float64), hyl.astype(np.float64)) + np.hypot(cxl.astype(np.float64), cyl.astype(np.float64))
    err4 = np.abs(t.astype(np.float64) - t_star)[keep]
    bud4 = (k["e1_raw"] * hp_true + k["d_fil"] / math.cos(k["th0"]) * nh64)[keep]
    assert (err4 <= bud4).all(), "(4) evaluation error exceeds e1 (|h'|+R) + d_fil |hd|: worst ratio %g" % (err4 / bud4).max()

    sign_in = np.signbit(t) & ~np.isnan(t)
    sure_chunk = ~(np.abs(t) < k["c1"] * nrm)
    sure_unit = ~(np.abs(t) < fma32(np.full(n, k["kappa2"], F), np.abs(p), (k["e1"] * nrm).astype(F)))
    bad = (keep & sure_chunk & (sign_in != verdict)).sum() + (keep & sure_unit & (sign_in != verdict)).sum()
    assert bad == 0, "(5) %d units decided wrongly outside the band" % bad
    # ---- (6) the direct form of k_refine_solve: hd = fl(h - c) itself, no chunk origin, band c1 |hd|_1
    pd = fma32(D, hdy, -(E * hdx).astype(F))
    td = (np.abs(pd) - fma32(G, hdx, (H * hdy).astype(F))).astype(F)
    sure_d = ~(np.abs(td) < (k["c1"] * (np.abs(hdx) + np.abs(hdy)).astype(F)).astype(F))
    sign_d = np.signbit(td) & ~np.isnan(td)
    bad_d = (keep & sure_d & (sign_d != verdict)).sum()
    assert bad_d == 0, "(6) %d units decided wrongly by the direct form" % bad_d
    return dict(n=int(keep.sum()), ref_err_u=float(err1.max() / U), dC_u=k["dC"] / U,
                eval_err=float((err4 / bud4).max()), uncertain=float((keep & ~sure_unit).mean()),
                flips=int((keep & (sign_in != verdict)).sum()))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
    rng = np.random.default_rng(20261017)
    print("thr      spread    units    max|ang-cos|/u (dC/u)   max eval err / budget   uncertain   sign flips inside the band")
    for thr in (0.5, 0.9, 0.97, 0.99, 0.999, 0.9999):
        for spread in (1e-2, 1e-4, 1e-5, 3e-6, 1e-6, 1e-7):
            r = run(F(thr), spread, n, rng)
            print("%-8g %-9g %-8d %6.2f (%5.2f)          %6.3f                  %8.5f    %d"
                  % (thr, spread, r["n"], r["ref_err_u"], r["dC_u"], r["eval_err"], r["uncertain"], r["flips"]))
    print("check_band ok: (1), (4), (5), (6) hold on every sample")


if __name__ == "__main__":
    main()
