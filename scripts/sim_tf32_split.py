"""CPU groundwork for the tensor-core version of the scoring loop: evaluation error of the
chunk-local test value t = |p| + s when the two linear forms are computed as 3xTF32 splits with K = 8
(D*hy and E*hx as hi*hi + hi*lo + lo*hi, the constant as hi + lo), and the stage-1 flag rate with the error bound
that this needs.  Development tool, no GPU; float64 is the ground truth, two accumulation models bracket the
(unspecified) tensor-core summation: 'ideal' = exact sum rounded once to float32, 'trunc' = every partial product and
every partial sum truncated to float32 in operand order.
usage: python scripts/sim_tf32_split.py   (imports oracle/ for the hypothesis generation: test infrastructure)"""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from casapose_b200 import synthetic  # noqa: E402
from oracle import philox_np  # noqa: E402
from oracle import ransac_voting_np as O  # noqa: E402

F = np.float32
u = 2.0 ** -24
thr = 0.99
th0 = math.acos(thr)
dC = 1.05 * (u + 8 * u * thr)
delta = 1.2 * (dC / math.sin(th0) + 8 * u)
k_lo, k_hi = math.tan(th0 - delta), math.tan(th0 + delta)
kap = 1 - 1 / (k_hi / k_lo)
e1_fp32 = 12 * 1.41421357 * u


def tf32(x):
    """round-to-nearest (ties away, like cvt.rna.tf32.f32) to 10 explicit mantissa bits"""
    b = np.asarray(x, F).view(np.uint32).astype(np.uint64)
    b = (b + np.uint64(0x1000)) & np.uint64(0xFFFFE000)
    return b.astype(np.uint32).view(F)


def split(x):
    hi = tf32(x)
    lo = tf32((np.asarray(x, F) - hi).astype(F))
    return hi.astype(np.float64), lo.astype(np.float64)


def trunc32(x):
    """float64 -> float32 rounded toward zero"""
    f = np.asarray(x, np.float64).astype(F)
    over = np.abs(f.astype(np.float64)) > np.abs(x)
    return np.where(over, np.nextafter(f, F(0)), f).astype(np.float64)


def form(terms, model):
    if model == "ideal":
        return np.asarray(sum(terms), np.float64).astype(F).astype(np.float64)
    acc = np.zeros_like(terms[0])
    for t in terms:
        acc = trunc32(acc + trunc32(t))
    return acc


def octn(x, y):
    ax, ay = np.abs(x), np.abs(y)
    return np.maximum(ax, ay) + 0.4142136 * np.minimum(ax, ay)


d = synthetic.make_frames(1, 480, 640, synthetic.CONFIG_8_IDS, variant="easy")
mask, vertex = d["mask"], d["vertex"]
worst = {"fp32": 0.0, "ideal": 0.0, "trunc": 0.0}
stash = []
for c in range(8):
    ys, xs = np.nonzero(mask[0, :, :, c])
    tn = len(ys)
    coords = np.stack([xs, ys], 1).astype(F) + F(0.5)
    direct = vertex[0, ys, xs][:, :, ::-1].astype(F)
    idx = philox_np.draw_idxs(1237, 0, c, 0, 512, 9, tn)
    hyp = O.generate_hypothesis(direct, coords, idx)
    for v in (0, 4, 8):
        dv = direct[:, v]
        q = (dv[:, 0] * dv[:, 0] + dv[:, 1] * dv[:, 1]).astype(F)
        inv = (F(1) / np.sqrt(q)).astype(F)
        D, E = (dv[:, 0] * inv).astype(F), (dv[:, 1] * inv).astype(F)
        G, H = (F(k_lo) * D).astype(F), (F(k_lo) * E).astype(F)
        h = hyp[:, v]
        ok = np.abs(h.sum(1)) > 1e-6
        for ch in range((tn + 127) // 128):
            sl = slice(ch * 128, min(tn, ch * 128 + 128))
            cx, cy = coords[sl, 0], coords[sl, 1]
            ox, oy = F(0.5) * (cx.min() + cx.max()), F(0.5) * (cy.min() + cy.max())
            cxl, cyl = (cx - ox).astype(F), (cy - oy).astype(F)
            rr = octn(cxl, cyl).max()
            # coefficients exactly as make_local_coef forms them (float32)
            P0 = (D[sl] * cyl - E[sl] * cxl).astype(F)
            A0 = (G[sl] * cxl + H[sl] * cyl).astype(F)
            hx, hy = (h[:, 0] - ox).astype(F), (h[:, 1] - oy).astype(F)
            Dd, Ed, Gd, Hd = (a.astype(np.float64)[None] for a in (D[sl], E[sl], G[sl], H[sl]))
            P0d, A0d = P0.astype(np.float64)[None], A0.astype(np.float64)[None]
            hxd, hyd = hx.astype(np.float64)[:, None], hy.astype(np.float64)[:, None]
            p_true = Dd * hyd - Ed * hxd - P0d
            s_true = A0d - Gd * hxd - Hd * hyd
            t_true = np.abs(p_true) + s_true
            scale = (np.abs(hxd) + np.abs(hyd)) + (np.abs(cxl) + np.abs(cyl)).astype(np.float64)[None]
            # shipped FP32 chain
            inner = (np.multiply(-E[sl][None].astype(np.float64), hxd) - P0d).astype(F)  # fma rounds once
            p32 = (Dd * hyd + inner.astype(np.float64)).astype(F)
            inner = (-Hd * hyd + A0d).astype(F)
            s32 = (-Gd * hxd + inner.astype(np.float64)).astype(F)
            t32 = (np.abs(p32) + s32).astype(F)
            worst["fp32"] = max(worst["fp32"], float((np.abs(t32.astype(np.float64) - t_true) / scale)[ok].max()))
            res = {}
            for model in ("ideal", "trunc"):
                Dh, Dl = split(D[sl]); Eh, El = split(E[sl]); Gh, Gl = split(G[sl]); Hh, Hl = split(H[sl])
                Ph, Pl = split(P0); Ah, Al = split(A0)
                xh, xl = split(hx); yh, yl = split(hy)
                xh, xl, yh, yl = xh[:, None], xl[:, None], yh[:, None], yl[:, None]
                p = form([Dh[None] * yh, Dh[None] * yl, Dl[None] * yh, -Eh[None] * xh, -Eh[None] * xl, -El[None] * xh,
                          -Ph[None] + 0 * xh, -Pl[None] + 0 * xh], model)
                s = form([Ah[None] + 0 * xh, Al[None] + 0 * xh, -Gh[None] * xh, -Gh[None] * xl, -Gl[None] * xh,
                          -Hh[None] * yh, -Hh[None] * yl, -Hl[None] * yh], model)
                t = (np.abs(p) + s).astype(F).astype(np.float64)
                worst[model] = max(worst[model], float((np.abs(t - t_true) / scale)[ok].max()))
                res[model] = np.abs(t).min(1)
            stash.append((np.abs(t32.astype(np.float64)).min(1), res["trunc"], octn(hx, hy).astype(np.float64), rr, ok))

print("max |t_eval - t_true| / (|h'|_1 + |c'|_1) in units of u = 2^-24:")
for k, v in worst.items():
    print("  %-6s %.2f u" % (k, v / u))
print("shipped bound: 6 u (doubled to 12*sqrt2 u = e1 = %.3g)" % e1_fp32)
for name, e1 in (("fp32 path, e1 = 12 sqrt2 u", e1_fp32), ("tensor path, e1 = 2 sqrt2 * (worst trunc, x4 margin)", 2 * 1.41421357 * 4 * worst["trunc"])):
    c1 = k_hi * kap + e1
    fl = n = 0
    for mn32, mntc, nh, rr, ok in stash:
        mn = mn32 if name.startswith("fp32") else mntc
        perm = np.arange(512).reshape(-1, 2, 32)
        ia, ib = perm[:, 0].ravel(), perm[:, 1].ravel()
        m = np.minimum(np.where(ok[ia], mn[ia], np.inf), np.where(ok[ib], mn[ib], np.inf))
        b = c1 * (np.maximum(np.where(ok[ia], nh[ia], 0), np.where(ok[ib], nh[ib], 0)) + rr)
        fl += int((m < b).sum())
        n += 256
    print("%-55s c1 = %.3g  flagged pairs = %.2f %%" % (name, c1, 100.0 * fl / n))
