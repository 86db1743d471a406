// K7 — batched PnP on the GPU (the step right after the voting path).
//
// Stands in for the reference's host-side pnp()
// (/root/reference/casapose/pose_estimation/ransac_voting.py:13-57: cv2.solvePnPRansac(EPNP) as the initial
// guess, then cv2.solvePnP(ITERATIVE) on ALL points, flip if t_z < 0, zeros on failure) and its callers
// map_offsets / map_pnp (:487-514).  OpenCV's RANSAC draws random minimal sets, so it cannot be replayed
// bit for bit; what is reproduced is its structure — a robust initial pose from point subsets, then the
// Levenberg-Marquardt minimum of the reprojection error over all points, which is what the reference returns:
//   * candidates: the full point set, every leave-one-out and every leave-two-out subset (46 for 9 points);
//     each is solved by a normalised DLT (reduced to a 4x4 symmetric eigen-problem), projected on SO(3) by a polar
//     iteration and polished by 5 LM steps; the candidate with the most points inside 12 px (cv2's
//     reprojectionError), then the smallest inlier cost, wins;
//   * final: LM over all points from the winner, float64 throughout.
// One warp per object, lanes = candidates.  oracle/pnp_np.py is the numpy restatement of this algorithm;
// tests compare both with the reference's cv2 sequence (identical ADD / ADD-S verdicts on the synthetic set).
#pragma once
#include "common.cuh"

namespace casa {

struct PnpParams {
  int n, vn;
  float reproj_px;  // 12: cv2.solvePnPRansac reprojectionError (:35)
};

__device__ __forceinline__ void mat3_inv_t(const double* R, double* out) {  // out = inverse(R)^T
  const double c00 = R[4] * R[8] - R[5] * R[7], c01 = R[5] * R[6] - R[3] * R[8], c02 = R[3] * R[7] - R[4] * R[6];
  const double c10 = R[2] * R[7] - R[1] * R[8], c11 = R[0] * R[8] - R[2] * R[6], c12 = R[1] * R[6] - R[0] * R[7];
  const double c20 = R[1] * R[5] - R[2] * R[4], c21 = R[2] * R[3] - R[0] * R[5], c22 = R[0] * R[4] - R[1] * R[3];
  const double det = R[0] * c00 + R[1] * c01 + R[2] * c02, id = 1.0 / det;
  out[0] = c00 * id; out[1] = c01 * id; out[2] = c02 * id;
  out[3] = c10 * id; out[4] = c11 * id; out[5] = c12 * id;
  out[6] = c20 * id; out[7] = c21 * id; out[8] = c22 * id;
}

__device__ __forceinline__ double mat3_det(const double* R) {
  return R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
}

// exp([w]x) (Rodrigues)
__device__ __forceinline__ void so3_exp(const double* w, double* E) {
  const double th = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  double a, b;  // E = I + a K + b K^2 with K = [w]x
  if (th < 1e-12) {
    a = 1.0;
    b = 0.5;
  } else {
    a = sin(th) / th;
    b = (1.0 - cos(th)) / (th * th);
  }
  const double x = w[0], y = w[1], z = w[2];
  E[0] = 1.0 - b * (y * y + z * z); E[1] = -a * z + b * x * y;       E[2] = a * y + b * x * z;
  E[3] = a * z + b * x * y;         E[4] = 1.0 - b * (x * x + z * z); E[5] = -a * x + b * y * z;
  E[6] = -a * y + b * x * z;        E[7] = a * x + b * y * z;         E[8] = 1.0 - b * (x * x + y * y);
}

// smallest-eigenvalue eigenvector of the symmetric 4x4 matrix A (destroyed), cyclic Jacobi
__device__ __forceinline__ void jacobi4_smallest(double (&A)[4][4], double (&vec)[4]) {
  double V[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0, diag = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      diag += A[i][i] * A[i][i];
#pragma unroll
      for (int j = i + 1; j < 4; ++j) off += A[i][j] * A[i][j];
    }
    if (off <= 1e-30 * diag || off == 0.0) break;
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int q = p + 1; q < 4; ++q) {
        const double apq = A[p][q];
        if (fabs(apq) < 1e-300) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // columns p, q
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // rows p, q
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int best = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (A[i][i] < A[best][best]) best = i;
#pragma unroll
  for (int k = 0; k < 4; ++k) vec[k] = best == 0 ? V[k][0] : best == 1 ? V[k][1] : best == 2 ? V[k][2] : V[k][3];
}

// normalised DLT on the points whose bit is set in `sel`; X [vn][3] object points, xn [vn][2] normalised image points.
// The 12 unknowns (rows p1, p2, p3 of the 3x4 projection) are reduced to the last row: for fixed p3 the optimum is
// p1 = S^-1 Sx p3, p2 = S^-1 Sy p3 with S = sum P P^T, Sx = sum x P P^T, Sy = sum y P P^T (P = normalised
// homogeneous point), and p3 is the smallest eigenvector of the 4x4 Schur complement
// M = sum (x^2 + y^2) P P^T - Sx S^-1 Sx - Sy S^-1 Sy.
__device__ bool pnp_dlt(const double* X, const double* xn, int vn, unsigned sel, double* R, double* t) {
  double c[3] = {0, 0, 0};
  int m = 0;
  for (int i = 0; i < vn; ++i)
    if ((sel >> i) & 1u) {
      c[0] += X[3 * i]; c[1] += X[3 * i + 1]; c[2] += X[3 * i + 2];
      ++m;
    }
  if (m < 6) return false;
  c[0] /= m; c[1] /= m; c[2] /= m;
  double ms = 0;
  for (int i = 0; i < vn; ++i)
    if ((sel >> i) & 1u) {
      const double a = X[3 * i] - c[0], b = X[3 * i + 1] - c[1], d = X[3 * i + 2] - c[2];
      ms += a * a + b * b + d * d;
    }
  if (!(ms > 0)) return false;
  const double s = 1.0 / sqrt(ms / m);
  double S[4][4], Sx[4][4], Sy[4][4], Sq[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) S[a][b] = Sx[a][b] = Sy[a][b] = Sq[a][b] = 0.0;
  for (int i = 0; i < vn; ++i) {
    if (!((sel >> i) & 1u)) continue;
    const double P[4] = {(X[3 * i] - c[0]) * s, (X[3 * i + 1] - c[1]) * s, (X[3 * i + 2] - c[2]) * s, 1.0};
    const double x = xn[2 * i], y = xn[2 * i + 1], q = x * x + y * y;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = a; b < 4; ++b) {
        const double pp = P[a] * P[b];
        S[a][b] += pp;
        Sx[a][b] += x * pp;
        Sy[a][b] += y * pp;
        Sq[a][b] += q * pp;
      }
  }
#pragma unroll
  for (int a = 1; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < a; ++b) {
      S[a][b] = S[b][a]; Sx[a][b] = Sx[b][a]; Sy[a][b] = Sy[b][a]; Sq[a][b] = Sq[b][a];
    }
  // Cholesky S = L L^T (S is positive definite for non-coplanar points)
  double L[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double d = S[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
    if (!(d > 1e-14 * S[j][j])) return false;
    L[j][j] = sqrt(d);
#pragma unroll
    for (int i = j + 1; i < 4; ++i) {
      double v = S[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k];
      L[i][j] = v / L[j][j];
    }
  }
  // Tx = S^-1 Sx, Ty = S^-1 Sy (column by column), M = Sq - Sx Tx - Sy Ty
  double Tx[4][4], Ty[4][4], M4[4][4];
#pragma unroll
  for (int col = 0; col < 4; ++col) {
    double zx[4], zy[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double vx = Sx[i][col], vy = Sy[i][col];
#pragma unroll
      for (int k = 0; k < i; ++k) {
        vx -= L[i][k] * zx[k];
        vy -= L[i][k] * zy[k];
      }
      zx[i] = vx / L[i][i];
      zy[i] = vy / L[i][i];
    }
#pragma unroll
    for (int i = 3; i >= 0; --i) {
      double vx = zx[i], vy = zy[i];
#pragma unroll
      for (int k = i + 1; k < 4; ++k) {
        vx -= L[k][i] * Tx[k][col];
        vy -= L[k][i] * Ty[k][col];
      }
      Tx[i][col] = vx / L[i][i];
      Ty[i][col] = vy / L[i][i];
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      double v = Sq[a][b];
#pragma unroll
      for (int k = 0; k < 4; ++k) v -= Sx[a][k] * Tx[k][b] + Sy[a][k] * Ty[k][b];
      M4[a][b] = v;
    }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = a + 1; b < 4; ++b) M4[a][b] = M4[b][a] = 0.5 * (M4[a][b] + M4[b][a]);
  double p3[4];
  jacobi4_smallest(M4, p3);
  double p[12];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    double v1 = 0, v2 = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v1 += Tx[a][k] * p3[k];
      v2 += Ty[a][k] * p3[k];
    }
    p[a] = v1;
    p[4 + a] = v2;
    p[8 + a] = p3[a];
  }
  double M[9], p4[3];
  for (int r = 0; r < 3; ++r) {
    for (int k = 0; k < 3; ++k) M[3 * r + k] = p[4 * r + k] * s;
    p4[r] = p[4 * r + 3] - (M[3 * r] * c[0] + M[3 * r + 1] * c[1] + M[3 * r + 2] * c[2]);
  }
  double det = mat3_det(M);
  if (det < 0) {
    for (int k = 0; k < 9; ++k) M[k] = -M[k];
    for (int k = 0; k < 3; ++k) p4[k] = -p4[k];
    det = -det;
  }
  if (!(det > 0) || !isfinite(det)) return false;
  const double sc = cbrt(det);
  for (int k = 0; k < 9; ++k) R[k] = M[k] / sc;
  for (int it = 0; it < 30; ++it) {  // polar iteration: R <- (R + R^-T) / 2
    double Ri[9], dmax = 0;
    mat3_inv_t(R, Ri);
    for (int k = 0; k < 9; ++k) {
      const double rn = 0.5 * (R[k] + Ri[k]);
      dmax = fmax(dmax, fabs(rn - R[k]));
      R[k] = rn;
    }
    if (dmax < 1e-15) break;
  }
  for (int k = 0; k < 3; ++k) t[k] = p4[k] / sc;
  return isfinite(t[0]) && isfinite(t[1]) && isfinite(t[2]) && isfinite(R[0]);
}

// squared reprojection cost of the selected points; uv in pixels
__device__ __forceinline__ double pnp_cost(const double* X, const double* uv, int vn, unsigned sel, const double* K4,
                                           const double* R, const double* t) {
  double cost = 0;
  for (int i = 0; i < vn; ++i) {
    if (!((sel >> i) & 1u)) continue;
    const double x = R[0] * X[3 * i] + R[1] * X[3 * i + 1] + R[2] * X[3 * i + 2] + t[0];
    const double y = R[3] * X[3 * i] + R[4] * X[3 * i + 1] + R[5] * X[3 * i + 2] + t[1];
    const double z = R[6] * X[3 * i] + R[7] * X[3 * i + 1] + R[8] * X[3 * i + 2] + t[2];
    const double ru = K4[0] * x / z + K4[2] - uv[2 * i], rv = K4[1] * y / z + K4[3] - uv[2 * i + 1];
    cost += ru * ru + rv * rv;
  }
  return cost;
}

// Levenberg-Marquardt on (rotation, translation), left-multiplicative rotation update; K4 = (fx, fy, cx, cy)
__device__ void pnp_lm(const double* X, const double* uv, int vn, unsigned sel, const double* K4, double* R, double* t, int iters) {
  double lam = 1e-3;
  double cost = pnp_cost(X, uv, vn, sel, K4, R, t);
  for (int it = 0; it < iters; ++it) {
    double H[36], g[6];
    for (int k = 0; k < 36; ++k) H[k] = 0;
    for (int k = 0; k < 6; ++k) g[k] = 0;
    for (int i = 0; i < vn; ++i) {
      if (!((sel >> i) & 1u)) continue;
      const double rx = R[0] * X[3 * i] + R[1] * X[3 * i + 1] + R[2] * X[3 * i + 2];
      const double ry = R[3] * X[3 * i] + R[4] * X[3 * i + 1] + R[5] * X[3 * i + 2];
      const double rz = R[6] * X[3 * i] + R[7] * X[3 * i + 1] + R[8] * X[3 * i + 2];
      const double x = rx + t[0], y = ry + t[1], z = rz + t[2];
      const double iz = 1.0 / z;
      const double ru = K4[0] * x * iz + K4[2] - uv[2 * i], rv = K4[1] * y * iz + K4[3] - uv[2 * i + 1];
      // d proj / d Xc
      const double a0 = K4[0] * iz, a2 = -K4[0] * x * iz * iz, b1 = K4[1] * iz, b2 = -K4[1] * y * iz * iz;
      // d Xc / d w = -[RX]x ; columns: (0, -rz, ry), (rz, 0, -rx), (-ry, rx, 0)   (= e_j x RX)
      double Ju[6], Jv[6];
      Ju[0] = a2 * ry;              Ju[1] = a0 * rz - a2 * rx; Ju[2] = -a0 * ry;
      Jv[0] = -b1 * rz + b2 * ry;   Jv[1] = -b2 * rx;          Jv[2] = b1 * rx;
      Ju[3] = a0; Ju[4] = 0;  Ju[5] = a2;
      Jv[3] = 0;  Jv[4] = b1; Jv[5] = b2;
      for (int a = 0; a < 6; ++a) {
        g[a] += Ju[a] * ru + Jv[a] * rv;
        for (int b = a; b < 6; ++b) H[a * 6 + b] += Ju[a] * Ju[b] + Jv[a] * Jv[b];
      }
    }
    for (int a = 0; a < 6; ++a)
      for (int b = 0; b < a; ++b) H[a * 6 + b] = H[b * 6 + a];
    bool ok = false;
    double dmax = 0, dc = 0;
    for (int tries = 0; tries < 10 && !ok; ++tries) {
      double Lm[36], d[6];
      bool spd = true;
      for (int k = 0; k < 36; ++k) Lm[k] = H[k];
      for (int k = 0; k < 6; ++k) Lm[k * 6 + k] += lam * H[k * 6 + k];
      for (int j = 0; j < 6 && spd; ++j) {  // Cholesky, lower triangle in place
        double s = Lm[j * 6 + j];
        for (int k = 0; k < j; ++k) s -= Lm[j * 6 + k] * Lm[j * 6 + k];
        if (!(s > 0)) {
          spd = false;
          break;
        }
        const double l = sqrt(s);
        Lm[j * 6 + j] = l;
        for (int i = j + 1; i < 6; ++i) {
          double v = Lm[i * 6 + j];
          for (int k = 0; k < j; ++k) v -= Lm[i * 6 + k] * Lm[j * 6 + k];
          Lm[i * 6 + j] = v / l;
        }
      }
      if (!spd) {
        lam *= 10;
        continue;
      }
      for (int i = 0; i < 6; ++i) {  // L y = -g
        double v = -g[i];
        for (int k = 0; k < i; ++k) v -= Lm[i * 6 + k] * d[k];
        d[i] = v / Lm[i * 6 + i];
      }
      for (int i = 5; i >= 0; --i) {  // L^T d = y
        double v = d[i];
        for (int k = i + 1; k < 6; ++k) v -= Lm[k * 6 + i] * d[k];
        d[i] = v / Lm[i * 6 + i];
      }
      double E[9], R2[9], t2[3];
      so3_exp(d, E);
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) R2[3 * r + c] = E[3 * r] * R[c] + E[3 * r + 1] * R[3 + c] + E[3 * r + 2] * R[6 + c];
      for (int k = 0; k < 3; ++k) t2[k] = t[k] + d[3 + k];
      const double c2 = pnp_cost(X, uv, vn, sel, K4, R2, t2);
      if (c2 < cost) {
        for (int k = 0; k < 9; ++k) R[k] = R2[k];
        for (int k = 0; k < 3; ++k) t[k] = t2[k];
        dc = cost - c2;
        cost = c2;
        lam = fmax(lam / 10, 1e-12);
        dmax = 0;
        for (int k = 0; k < 6; ++k) dmax = fmax(dmax, fabs(d[k]));
        ok = true;
      } else {
        lam *= 10;
      }
    }
    if (!ok || dmax < 1e-12 || dc < 1e-14 * fmax(cost, 1e-30)) break;
  }
}

// transform_points_back_tf (:92-121) in float32, applied to one point
__device__ __forceinline__ float2 pnp_unmap(float2 p, const float* o) {
  // offsets: [0]=h_crop, [1]=w_crop, [8]=sx, [9]=sy, [4]=dx, [5]=dy, [6]=angle, [7]=scale   (:494-504)
  const float sx = o[8], sy = o[9];
  float x = __fadd_rn(__fdiv_rn(p.x, o[7]), o[1]);
  float y = __fadd_rn(__fdiv_rn(p.y, o[7]), o[0]);
  x = __fsub_rn(x, o[4]);
  y = __fsub_rn(y, o[5]);
  const float ang = __fmul_rn(-o[6], 0.017453292519943295f);
  const float a = cosf(ang), b = sinf(ang);
  const float cx = __fdiv_rn(sx, 2.0f), cy = __fdiv_rn(sy, 2.0f);
  const float c = __fsub_rn(__fmul_rn(__fsub_rn(1.0f, a), cx), __fmul_rn(b, cy));
  const float d = __fadd_rn(__fmul_rn(b, cx), __fmul_rn(__fsub_rn(1.0f, a), cy));
  return make_float2(__fadd_rn(__fadd_rn(__fmul_rn(a, x), __fmul_rn(b, y)), c),
                     __fadd_rn(__fadd_rn(__fmul_rn(-b, x), __fmul_rn(a, y)), d));
}

// one warp per object.  pts2d [n,vn,2] (x,y) px, pts3d [n,vn,3], cam [n,3,3] (zero skew), offsets [n,10] or NULL
__global__ void __launch_bounds__(128) k_pnp(PnpParams pp, const float* __restrict__ pts2d, const float* __restrict__ pts3d,
                                            const float* __restrict__ cam, const float* __restrict__ offsets,
                                            float* __restrict__ poses) {
  const int obj = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (obj >= pp.n) return;
  const int vn = pp.vn;
  double X[48], uv[32], xn[32];
  float sum = 0.f;
  for (int i = 0; i < vn; ++i) {
    float2 p = make_float2(pts2d[((size_t)obj * vn + i) * 2], pts2d[((size_t)obj * vn + i) * 2 + 1]);
    sum = __fadd_rn(sum, __fadd_rn(p.x, p.y));
    uv[2 * i] = p.x;
    uv[2 * i + 1] = p.y;
  }
  float* out = poses + (size_t)obj * 12;
  bool dead = fabsf(sum) < 0.01f;  // map_offsets :489 / map_pnp :510
  if (!dead && offsets) {
    sum = 0.f;
    for (int i = 0; i < vn; ++i) {
      const float2 q = pnp_unmap(make_float2((float)uv[2 * i], (float)uv[2 * i + 1]), offsets + (size_t)obj * 10);
      uv[2 * i] = q.x;
      uv[2 * i + 1] = q.y;
      sum = __fadd_rn(sum, __fadd_rn(q.x, q.y));
    }
    dead = fabsf(sum) < 0.01f;
  }
  if (dead) {
    if (lane < 12) out[lane] = 0.f;
    return;
  }
  const float* Kf = cam + (size_t)obj * 9;
  const double K4[4] = {Kf[0], Kf[4], Kf[2], Kf[5]};
  for (int i = 0; i < vn; ++i) {
    X[3 * i] = pts3d[((size_t)obj * vn + i) * 3];
    X[3 * i + 1] = pts3d[((size_t)obj * vn + i) * 3 + 1];
    X[3 * i + 2] = pts3d[((size_t)obj * vn + i) * 3 + 2];
    xn[2 * i] = (uv[2 * i] - K4[2]) / K4[0];
    xn[2 * i + 1] = (uv[2 * i + 1] - K4[3]) / K4[1];
  }
  // candidates: 0 = all points, 1..vn = leave one out, then leave two out (a < b)
  const unsigned full = vn >= 32 ? 0xffffffffu : ((1u << vn) - 1u);
  const int ncand = 1 + vn + vn * (vn - 1) / 2;
  double bestR[9], bestT[3], bestCost = 1e300;
  int bestInl = -1, bestIdx = 0x7fffffff;
  for (int cnd = lane; cnd < ncand; cnd += 32) {
    unsigned sel = full;
    if (cnd >= 1 && cnd <= vn) {
      sel &= ~(1u << (cnd - 1));
    } else if (cnd > vn) {
      int k = cnd - vn - 1, a = 0;
      while (k >= vn - 1 - a) {
        k -= vn - 1 - a;
        ++a;
      }
      sel &= ~(1u << a);
      sel &= ~(1u << (a + 1 + k));
    }
    double R[9], t[3];
    if (!pnp_dlt(X, xn, vn, sel, R, t)) continue;
    pnp_lm(X, uv, vn, sel, K4, R, t, 5);
    int ninl = 0;
    double cost = 0;
    bool front = true;
    for (int i = 0; i < vn; ++i) {
      const double x = R[0] * X[3 * i] + R[1] * X[3 * i + 1] + R[2] * X[3 * i + 2] + t[0];
      const double y = R[3] * X[3 * i] + R[4] * X[3 * i + 1] + R[5] * X[3 * i + 2] + t[1];
      const double z = R[6] * X[3 * i] + R[7] * X[3 * i + 1] + R[8] * X[3 * i + 2] + t[2];
      if (!(z > 0)) front = false;
      const double ru = K4[0] * x / z + K4[2] - uv[2 * i], rv = K4[1] * y / z + K4[3] - uv[2 * i + 1];
      const double e2 = ru * ru + rv * rv;
      if (e2 < (double)pp.reproj_px * pp.reproj_px) {
        ++ninl;
        cost += e2;
      }
    }
    if (!front || !isfinite(cost)) continue;
    if (ninl > bestInl || (ninl == bestInl && cost < bestCost)) {  // candidates come in ascending index per lane
      bestInl = ninl;
      bestCost = cost;
      bestIdx = cnd;
      for (int k = 0; k < 9; ++k) bestR[k] = R[k];
      for (int k = 0; k < 3; ++k) bestT[k] = t[k];
    }
  }
  // warp arg-max of (inliers, -cost, -index)
  int winner = lane;
  {
    int inl = bestInl, idx = bestIdx, who = lane;
    double cst = bestCost;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const int oi = __shfl_xor_sync(0xffffffffu, inl, o), ox = __shfl_xor_sync(0xffffffffu, idx, o);
      const int ow = __shfl_xor_sync(0xffffffffu, who, o);
      const double oc = __shfl_xor_sync(0xffffffffu, cst, o);
      const bool better = oi > inl || (oi == inl && (oc < cst || (oc == cst && ox < idx)));
      if (better) {
        inl = oi;
        idx = ox;
        who = ow;
        cst = oc;
      }
    }
    winner = who;
    bestInl = inl;
  }
  if (bestInl < 0) {  // no candidate produced a pose in front of the camera: the reference returns zeros on failure
    if (lane < 12) out[lane] = 0.f;
    return;
  }
  if (lane == winner) {
    pnp_lm(X, uv, vn, full, K4, bestR, bestT, 50);  // cv2.solvePnP(ITERATIVE, useExtrinsicGuess) on all points (:37-46)
    bool fin = true;
    for (int k = 0; k < 3; ++k) fin = fin && isfinite(bestT[k]);
    for (int k = 0; k < 9; ++k) fin = fin && isfinite(bestR[k]);
    const double sg = bestT[2] < 0 ? -1.0 : 1.0;  // :53-55
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) out[4 * r + c] = fin ? (float)(sg * bestR[3 * r + c]) : 0.f;
      out[4 * r + 3] = fin ? (float)(sg * bestT[r]) : 0.f;
    }
  }
}

}  // namespace casa
