// The inlier predicate of voting_for_hypothesis
// (/root/reference/casapose/pose_estimation/ransac_voting.py:230-249) in two forms:
//
//  * exact_inlier(): the reference's float32 op sequence, one IEEE rounding per TensorFlow
//    op (__f*_rn intrinsics are never contracted into FMAs).  This DEFINES the result.
//
//  * the filtered predicate (chunk-local form, what k_score evaluates): per (pixel, keypoint) the unit
//    direction d^ = d/|d| is turned into two linear forms of hd = h - c,
//        p = d^ x hd            (= |hd| sin(theta))
//        s = -k (d^ . hd)       (= -k |hd| cos(theta)),      k = tan(theta0),  theta0 = acos(inlier_thresh)
//    so that  t = |p| + s = |hd| sin(theta - theta0) / cos(theta0)  is negative exactly for theta < theta0.
//    The band is CENTRED on the threshold: sign(t) is the reference's verdict whenever
//        |t| >= w |hd| + E,        w = sin(delta) / cos(theta0)
//    where delta covers (i) the worst-case rounding of the reference's own sequence around the true
//    cosine and (ii) the rounding of the filter's coefficients, and E is the evaluation error of t
//    (derivation: filter_consts() below and DESIGN.md "Filtered predicate"; executable check:
//    scripts/check_band.py).  Units inside the band (~2e-5 of all units) are re-evaluated with
//    exact_inlier().  Vote counts are therefore bit-identical to the exact sequence while the common path
//    costs 5 FP32-pipe instructions per unit instead of ~30.
#pragma once
#include <math.h>

#include "common.cuh"

namespace casa {

constexpr float kEps1e6 = 1e-6f;          // python 1e-6 converted to float32 (:226, :240, :242)
constexpr float kDirMax = 1073741824.0f;  // 2^30: direction components above this take the exact path
constexpr float kHypMax = 1.152921504606846976e18f;  // 2^60: hypotheses beyond this take the exact path
constexpr float kNearCentre = 8e-6f;      // hypotheses this close to a pixel centre take the exact path

// ---------------------------------------------------------------- exact (reference) sequence

// tf.norm(direct, axis=-1)  (:238)
__device__ __forceinline__ float exact_norm(float x, float y) {
  return __fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));
}

// voting_for_hypothesis for one unit (:236-247). (dx,dy) = direct, nd = exact_norm(dx,dy).
__device__ __forceinline__ bool exact_inlier(float hx, float hy, float cx, float cy, float dx, float dy,
                                             float nd, float thr) {
  const float hdx = __fsub_rn(hx, cx);
  const float hdy = __fsub_rn(hy, cy);
  const float nh = exact_norm(hdx, hdy);
  const bool valid = (nd > kEps1e6) && (nh > kEps1e6) && (fabsf(__fadd_rn(hx, hy)) > kEps1e6);
  const float dot = __fadd_rn(__fmul_rn(dx, hdx), __fmul_rn(dy, hdy));
  const float ang = __fdiv_rn(dot, __fmul_rn(nd, nh));
  return valid && (ang > thr);
}

// generate_hypothesis for one (h, v) (:219-226).  c0/c1 = coords of the pair, d0/d1 = (dx,dy).
__device__ __forceinline__ float2 exact_hypothesis(float2 c0, float2 c1, float2 d0, float2 d1) {
  const float det = __fsub_rn(__fmul_rn(d1.x, d0.y), __fmul_rn(d1.y, d0.x));
  const float num = __fsub_rn(__fmul_rn(__fsub_rn(c1.y, c0.y), d1.x), __fmul_rn(__fsub_rn(c1.x, c0.x), d1.y));
  const float u = __fdiv_rn(num, det);
  float2 hp;
  hp.x = __fadd_rn(c0.x, __fmul_rn(d0.x, u));
  hp.y = __fadd_rn(c0.y, __fmul_rn(d0.y, u));
  if (!(fabsf(det) > kEps1e6)) hp = make_float2(0.f, 0.f);
  return hp;
}

// ---------------------------------------------------------------- filtered predicate

struct PixCoef {
  float D, E;  // d^ = (dx, dy) / |d|
  float G, H;  // k * d^
};

// float32 bit pattern of the smallest q with sqrt_rn(q) > 1e-6f (sqrt_rn is monotone), so that
//   tf.norm(direct) > 1e-6  (:238-240)   <=>   fl(fl(dx*dx) + fl(dy*dy)) >= kNormSqMin
// without evaluating the square root (checked on the host: tests/test_oracle_ransac.py).
constexpr uint32_t kNormSqMinBits = 0x2b8cbcceu;  // 1.0000002e-12f

// Returns false if this (pixel, keypoint) cannot use the filter (non-finite or huge direction).
__device__ __forceinline__ bool make_coef(float dx, float dy, float k, PixCoef& c) {
  c.D = c.E = c.G = c.H = 0.f;  // zero forms: never an inlier
  const float ax = fabsf(dx), ay = fabsf(dy);
  if (!(ax <= kDirMax && ay <= kDirMax)) return false;  // also catches NaN / inf
  const float q = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
  if (q >= __uint_as_float(kNormSqMinBits)) {  // <=> exact_norm(dx,dy) > 1e-6f (:240)
    const float inv = rsqrtf(q);               // 2 ulp; a common scale of (D,E,G,H) cannot change a sign
    c.D = dx * inv;
    c.E = dy * inv;
    c.G = k * c.D;
    c.H = k * c.E;
  }
  return true;
}

// ---------------------------------------------------------------- chunk-local form (what k_score evaluates)
// upper bound of sqrt(x^2 + y^2):  max + (sqrt2 - 1) min  (exact at 0 and 45 degrees, concave in between)
__device__ __forceinline__ float oct_norm(float x, float y) {
  const float ax = fabsf(x), ay = fabsf(y);
  return fmaf(0.4142136f, fminf(ax, ay), fmaxf(ax, ay)) * 1.0000005f;
}

// Coefficients of one pixel relative to the chunk origin o: (cxl, cyl) = c - o (exact).
//   A = (D, -E, -P0, A0), B = (-G, -H);  invalid pixels (|d| <= 1e-6): t = +1e30 for every hypothesis.
// Returns false if the direction cannot use the filter (non-finite / huge).
__device__ __forceinline__ bool make_local_coef(float cxl, float cyl, float dx, float dy, float k, float4& A, float2& B) {
  A = make_float4(0.f, 0.f, 0.f, 1.0e30f);
  B = make_float2(0.f, 0.f);
  PixCoef pc;
  if (!make_coef(dx, dy, k, pc)) return false;
  if (pc.D != 0.f || pc.E != 0.f) {
    A = make_float4(pc.D, -pc.E, -(pc.D * cyl - pc.E * cxl), pc.G * cxl + pc.H * cyl);
    B = make_float2(-pc.G, -pc.H);
  }
  return true;
}

// One unit in the chunk-local form, (hxl, hyl) = fl(h - o):  t = |p| + s, inlier <=> sign bit of t.
__device__ __forceinline__ float local_unit(const float4& A, const float2& B, float hxl, float hyl, float& p) {
  p = fmaf(A.x, hyl, fmaf(A.y, hxl, A.z));
  return fabsf(p) + fmaf(B.x, hxl, fmaf(B.y, hyl, A.w));
}

// 0 = normal (filter), 1 = zero count by construction, 2 = exact list
__device__ __forceinline__ int classify_hypothesis(float hx, float hy, bool fast_ok) {
  if (!(fabsf(hx) <= 3.0e38f && fabsf(hy) <= 3.0e38f)) return 1;  // inf / NaN: every ang is NaN or 0/inf
  if (!(fabsf(__fadd_rn(hx, hy)) > kEps1e6)) return 1;             // :241-243 guard fails for every pixel
  if (!fast_ok) return 2;
  if (fabsf(hx) > kHypMax || fabsf(hy) > kHypMax) return 2;        // hd^2 may overflow in the reference
  bool nearx = false, neary = false;
  if (fabsf(hx) < 4194304.f) nearx = fabsf(hx - (floorf(hx) + 0.5f)) <= kNearCentre;
  if (fabsf(hy) < 4194304.f) neary = fabsf(hy - (floorf(hy) + 0.5f)) <= kNearCentre;
  if (nearx && neary) return 2;  // |hd| may be <= 1e-6 for one pixel (:240 norm_hyp guard)
  return 0;
}

// ---------------------------------------------------------------- host: filter constants

// Derivation (u = 2^-24; theta = angle between the stored direction d and the reference's hd = fl(h - c)):
//  (1) the reference's float32 sequence (:236-247) returns ang with |ang - cos(theta)| <= u + 7u|cos(theta)| + O(u^2):
//      the two products and the sum of `dot` give |d||hd| u (1 + |cos| + u), each norm carries 2u (two squares and
//      a sum of positive terms, halved by the root, plus the root's own rounding), their product u, the divide u.
//      dC = 1.05 (u + 8u thr) bounds that for every theta near the threshold; in angle d_ref = dC / sin(theta0).
//  (2) the filter's own direction (D,E) = fl(d * rsqrt(q)) and (G,H) = fl(k (D,E)) deviate from d^ by < 3u of angle
//      (component roundings; the common scale error of rsqrt cannot change a sign), k = fl(tan(theta0)) by < 1u:
//      d_fil = 8u.
//  (3) delta = 1.2 (d_ref + d_fil); cos(theta0 -+ delta) is checked against thr +- dC below, so every unit with
//      |theta - theta0| >= delta is decided by the reference exactly as sign(theta - theta0) says.
//  (4) real arithmetic: t* = |hd| sin(theta - theta0)/cos(theta0), so |theta - theta0| < delta  =>  |t*| < w |hd|,
//      w = sin(delta)/cos(theta0).  The evaluated t differs from t* by at most e (|h'| + R): 6u (|h'|_1 + |c'|_1)
//      from the two FMA chains, the final add and the coefficient roundings, 2 sqrt2 (1+k) u (|h'| + R) from
//      h' - c' = fl(h - o) - (c - o) standing in for fl(h - c);  e1 = 12 sqrt2 u covers both with a factor 1.45.
//  (5) chunk level: |hd| <= |h'| + R, hence  min|t| >= c1 (|h'| + R),  c1 = w + e1,  proves every sign of the chunk;
//      unit level (stage 2): inside the band |p*| >= |hd| sin(theta0 - delta), hence a unit can be uncertain only if
//      |t| < kappa2 |p| + e2 (|h'| + R),  kappa2 = w / sin(theta0 - delta),  e2 = (1 + kappa2) e1.
// scripts/check_band.py evaluates (1), (4) and (5) against float64 on adversarial grids; casa_selftest_filter
// does the same on the device against exact_inlier().
inline FilterConsts filter_consts(float inlier_thresh, int force_exact) {
  FilterConsts f;
  f.thr = inlier_thresh;
  f.k_mid = 0.f;
  f.c1 = 0.f;
  f.e1 = 0.f;
  f.kappa2 = 0.f;
  f.fast_ok = 0;
  const double thr = (double)inlier_thresh;
  if (force_exact || !(thr >= 0.5 && thr <= 0.9999)) return f;
  const double u = 5.9604644775390625e-8;  // 2^-24
  const double theta0 = acos(thr);
  const double s0 = sin(theta0);
  const double dC = 1.05 * (u + 8.0 * u * thr);
  const double d_ref = dC / s0;   // angular half-width from the reference's own rounding
  const double d_fil = 8.0 * u;   // angular error of the filter's two linear forms
  const double delta = 1.2 * (d_ref + d_fil);
  if (!(theta0 - delta > 1e-3)) return f;
  // verify the band really covers dC on both sides (curvature of cos)
  if (!(cos(theta0 - delta) >= thr + dC && cos(theta0 + delta) <= thr - dC)) return f;
  f.k_mid = (float)tan(theta0);
  const double w = sin(delta) / cos(theta0);
  const double kappa2 = w / sin(theta0 - delta);
  const double e1 = 12.0 * 1.41421357 * u;
  f.e1 = (float)(1.001 * (1.0 + kappa2) * e1);
  f.kappa2 = (float)(1.001 * kappa2);
  f.c1 = (float)(1.001 * (w + e1));
  f.fast_ok = 1;
  return f;
}

}  // namespace casa
