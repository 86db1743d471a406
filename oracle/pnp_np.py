"""numpy restatement of the batched GPU PnP (casapose_b200/csrc/pnp.cuh).

TEST INFRASTRUCTURE (see oracle/__init__.py).
The reference's pnp() (/root/reference/casapose/pose_estimation/ransac_voting.py:13-57) is
cv2.solvePnPRansac(EPNP) -> cv2.solvePnP(ITERATIVE) on all points.  OpenCV's RANSAC is randomised, so the GPU
kernel reproduces its structure rather than its bits; this file restates the kernel's algorithm step by step
(candidate subsets, normalised DLT, polar projection, LM) so that GPU-vs-oracle is a tight numerical check,
while tests/test_pnp.py compares both against the reference's actual cv2 sequence (oracle/pose_np.pnp)."""
import itertools

import numpy as np


def _polar(M):
    R = M.copy()
    for _ in range(30):
        Rn = 0.5 * (R + np.linalg.inv(R).T)
        done = np.abs(Rn - R).max() < 1e-15
        R = Rn
        if done:
            break
    return R


def dlt(X, xn):
    c = X.mean(0)
    s = 1.0 / np.sqrt(((X - c) ** 2).sum(1).mean())
    Xs = (X - c) * s
    P = np.concatenate([Xs, np.ones((len(X), 1))], 1)  # [m,4]
    x, y = xn[:, 0], xn[:, 1]
    PP = P[:, :, None] * P[:, None, :]
    S, Sx, Sy = PP.sum(0), (x[:, None, None] * PP).sum(0), (y[:, None, None] * PP).sum(0)
    Sq = ((x * x + y * y)[:, None, None] * PP).sum(0)
    np.linalg.cholesky(S)  # positive definite for non-coplanar points (LinAlgError otherwise)
    Tx, Ty = np.linalg.solve(S, Sx), np.linalg.solve(S, Sy)
    M4 = Sq - Sx @ Tx - Sy @ Ty  # Schur complement on the last row of the projection
    M4 = 0.5 * (M4 + M4.T)
    _, v = np.linalg.eigh(M4)
    p3 = v[:, 0]
    p = np.stack([Tx @ p3, Ty @ p3, p3])
    M = p[:, :3] * s
    p4 = p[:, 3] - M @ c
    if np.linalg.det(M) < 0:
        M, p4 = -M, -p4
    sc = np.cbrt(np.linalg.det(M))
    return _polar(M / sc), p4 / sc


def _skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def _exp(w):
    th = np.linalg.norm(w)
    K = _skew(w)
    if th < 1e-12:
        return np.eye(3) + K + 0.5 * K @ K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th**2 * K @ K


def lm(X, uv, K4, R, t, iters):
    fx, fy, cx, cy = K4

    def residual(R, t):
        Xc = X @ R.T + t
        return np.stack([fx * Xc[:, 0] / Xc[:, 2] + cx - uv[:, 0], fy * Xc[:, 1] / Xc[:, 2] + cy - uv[:, 1]], 1).ravel(), Xc

    lam = 1e-3
    r, Xc = residual(R, t)
    cost = r @ r
    for _ in range(iters):
        J = np.zeros((2 * len(X), 6))
        RX = X @ R.T
        for i in range(len(X)):
            x, y, z = Xc[i]
            dp = np.array([[fx / z, 0, -fx * x / z**2], [0, fy / z, -fy * y / z**2]])
            J[2 * i : 2 * i + 2, :3] = dp @ (-_skew(RX[i]))
            J[2 * i : 2 * i + 2, 3:] = dp
        H, g = J.T @ J, J.T @ r
        ok, dc, d = False, 0.0, np.zeros(6)
        for _ in range(10):
            try:
                L = np.linalg.cholesky(H + lam * np.diag(np.diag(H)))
            except np.linalg.LinAlgError:
                lam *= 10
                continue
            d = np.linalg.solve(L.T, np.linalg.solve(L, -g))
            R2, t2 = _exp(d[:3]) @ R, t + d[3:]
            r2, Xc2 = residual(R2, t2)
            c2 = r2 @ r2
            if c2 < cost:
                R, t, r, Xc, dc, cost = R2, t2, r2, Xc2, cost - c2, c2
                lam = max(lam / 10, 1e-12)
                ok = True
                break
            lam *= 10
        if not ok or np.abs(d).max() < 1e-12 or dc < 1e-14 * max(cost, 1e-30):
            break
    return R, t


def pnp(points_3d, points_2d, camera_matrix, reproj_px=12.0):
    """[vn,3], [vn,2] (x,y), [3,3] -> [3,4] float32, the algorithm of k_pnp."""
    uv32 = np.asarray(points_2d, np.float32)
    if abs(float(uv32.sum(dtype=np.float32))) < 0.01:
        return np.zeros((3, 4), np.float32)
    X = np.asarray(points_3d, np.float32).astype(np.float64)
    uv = uv32.astype(np.float64)
    K = np.asarray(camera_matrix, np.float32).astype(np.float64)
    K4 = (K[0, 0], K[1, 1], K[0, 2], K[1, 2])
    xn = np.stack([(uv[:, 0] - K4[2]) / K4[0], (uv[:, 1] - K4[3]) / K4[1]], 1)
    n = len(X)
    subsets = [list(range(n))] + [[i for i in range(n) if i != a] for a in range(n)]
    subsets += [[i for i in range(n) if i not in ab] for ab in itertools.combinations(range(n), 2)]
    best = None
    for idx in subsets:
        if len(idx) < 6:
            continue
        with np.errstate(all="ignore"):
            try:
                R, t = dlt(X[idx], xn[idx])
            except np.linalg.LinAlgError:
                continue
            if not (np.isfinite(t).all() and np.isfinite(R).all()):
                continue
            R, t = lm(X[idx], uv[idx], K4, R, t, 5)
            Xc = X @ R.T + t
            e2 = (K4[0] * Xc[:, 0] / Xc[:, 2] + K4[2] - uv[:, 0]) ** 2 + (K4[1] * Xc[:, 1] / Xc[:, 2] + K4[3] - uv[:, 1]) ** 2
        if not (Xc[:, 2] > 0).all():
            continue
        inl = e2 < reproj_px * reproj_px
        cost = float(e2[inl].sum())
        if not np.isfinite(cost):
            continue
        key = (int(inl.sum()), -cost)
        if best is None or key > best[0]:
            best = (key, R, t)
    if best is None:
        return np.zeros((3, 4), np.float32)
    R, t = lm(X, uv, K4, best[1], best[2], 50)
    if not (np.isfinite(t).all() and np.isfinite(R).all()):
        return np.zeros((3, 4), np.float32)
    if t[2] < 0:
        R, t = -R, -t
    return np.concatenate([R, t[:, None]], axis=1).astype(np.float32)
