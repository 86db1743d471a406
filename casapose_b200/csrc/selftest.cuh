// Device self-tests that ship with the library:
//   k_selftest_filter  adversarial check of the filtered predicate against the exact sequence
//   k_fma_peak         FP32 issue-rate micro-benchmark: the measured denominator of "% of FP32 peak"
#pragma once
#include "common.cuh"
#include "philox.cuh"
#include "predicate.cuh"

namespace casa {

__device__ __forceinline__ float u01(uint32_t w) { return (float)(w >> 8) * 5.9604644775390625e-8f; }

// out[0]=tested, out[1]=mismatches, out[2]=uncertain (sent to exact), out[3]=exact inliers
__global__ void __launch_bounds__(256) k_selftest_filter(unsigned long long n, uint32_t k0, uint32_t k1, FilterConsts fc,
                                                         double theta0, float spread, unsigned long long* out) {
  unsigned long long tested = 0, bad = 0, unc = 0, inl = 0;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint4 r0 = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)(i >> 32), 0u, 77u), k0, k1);
    const uint4 r1 = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)(i >> 32), 1u, 77u), k0, k1);
    // spread > 0: image-sized coordinates, |d| in 2^[-10,10], distances 2^[-6,14];
    // spread < 0 ("extreme"): coordinates up to 65535, |d| in 2^[-19,29], distances 2^[-17,59]
    const bool extreme = spread < 0.f;
    const float cx = (float)(r0.x % (extreme ? 65535u : 1920u)) + 0.5f, cy = (float)(r0.y % (extreme ? 65535u : 1080u)) + 0.5f;
    const double phi = (double)u01(r0.z) * 6.283185307179586;
    const float mag = extreme ? exp2f(u01(r0.w) * 48.f - 19.f) : exp2f(u01(r0.w) * 20.f - 10.f);
    const float dx = mag * (float)cos(phi), dy = mag * (float)sin(phi);
    const double sgn = (r1.x & 1u) ? 1.0 : -1.0;
    const double th = theta0 + (double)fabsf(spread) * (2.0 * (double)u01(r1.y) - 1.0);
    const double dist = extreme ? exp2((double)u01(r1.z) * 76.0 - 17.0) : exp2((double)u01(r1.z) * 20.0 - 6.0);
    // the direction actually stored is the rounded (dx,dy): aim relative to it
    const double phir = atan2((double)dy, (double)dx);
    const float hx = (float)((double)cx + dist * cos(phir + sgn * th));
    const float hy = (float)((double)cy + dist * sin(phir + sgn * th));
    if (classify_hypothesis(hx, hy, true) != 0) continue;
    PixCoef pc;
    if (!make_coef(cx, cy, dx, dy, fc.k_lo, pc)) continue;
    bool lo, hi;
    filter_test(pc, hx, hy, fc.kappa, lo, hi);
    const bool ex = exact_inlier(hx, hy, cx, cy, dx, dy, exact_norm(dx, dy), fc.thr);
    ++tested;
    bad += (lo && !ex) || (!hi && ex) || (lo && !hi);
    unc += (hi && !lo);
    inl += ex;
    // ---- the chunk-local form exactly as k_score evaluates it: random chunk origin within 90 px of the pixel
    // (one fractional bit, like a bounding-box centre), chunk radius >= this pixel's offset
    const float ox = cx + 0.5f * (float)((int)(r1.w % 361u) - 180), oy = cy + 0.5f * (float)((int)((r1.w >> 10) % 361u) - 180);
    const float cxl = cx - ox, cyl = cy - oy;
    float4 A;
    float2 B;
    if (!make_local_coef(cxl, cyl, dx, dy, fc.k_lo, A, B)) continue;
    const float hxl = hx - ox, hyl = hy - oy;
    float pl;
    const float t = local_unit(A, B, hxl, hyl, pl);
    const float rr = oct_norm(cxl, cyl) * (1.0f + u01(r1.x));
    const float nrm = oct_norm(hxl, hyl) + rr;
    const bool sign_in = (__float_as_uint(t) >> 31) != 0u;
    const bool sure_unit = !(fabsf(t) < fmaf(fc.kappa2, fabsf(pl), fc.e1 * nrm));   // what stage 2 relies on
    const bool sure_chunk = !(fabsf(t) < fc.c1 * nrm);                               // what the main loop relies on
    bad += (sure_unit && sign_in != ex) || (sure_chunk && sign_in != ex);
    unc += !sure_unit;
  }
  atomicAdd(&out[0], tested);
  atomicAdd(&out[1], bad);
  atomicAdd(&out[2], unc);
  atomicAdd(&out[3], inl);
}

// Scoring-loop instruction mixes for design exploration (8 hypotheses per lane, one pixel per step).
//   PACK: 0 scalar, 1 hd as FADD2, 2 hd as FADD2 and p as FMUL2+FFMA2
//   UNC : 0 thi FFMA + second LEA.HI counter, 1 w = |tlo| - kappa|p| + FMNMX3 per pair, 2 FMNMX3 of |tlo| per pair
template <int PACK, int UNC>
__device__ __forceinline__ unsigned mix_loop(int iters, float a, float b) {
  float2 hx2[4], hy2[4];
  unsigned nlo[8], nhi[8];
  float mn[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    hx2[i] = make_float2(a * (float)(2 * i + 1), a * (float)(2 * i + 2));
    hy2[i] = make_float2(b + (float)(2 * i), b + (float)(2 * i + 1));
    mn[i] = 3.0e38f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) nlo[i] = nhi[i] = 0u;
  float cx = a, cy = b;
  const float D = 0.6f, nE = -0.8f, nG = -0.08f, nH = -0.11f, nk = -1e-4f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const float2 ncx2 = make_float2(-cx, -cx), ncy2 = make_float2(-cy, -cy);
      const float2 D2 = make_float2(D + cx, D + cx), nE2 = make_float2(nE + cy, nE + cy);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 hdx, hdy, pv;
        if (PACK >= 1) {
          hdx = __fadd2_rn(hx2[i], ncx2);
          hdy = __fadd2_rn(hy2[i], ncy2);
        } else {
          hdx = make_float2(hx2[i].x - cx, hx2[i].y - cx);
          hdy = make_float2(hy2[i].x - cy, hy2[i].y - cy);
        }
        if (PACK >= 2) {
          pv = __ffma2_rn(D2, hdy, __fmul2_rn(nE2, hdx));
        } else {
          pv.x = fmaf(D2.x, hdy.x, __fmul_rn(nE2.x, hdx.x));
          pv.y = fmaf(D2.x, hdy.y, __fmul_rn(nE2.x, hdx.y));
        }
        const float tla = fmaf(nG, hdx.x, fmaf(nH, hdy.x, fabsf(pv.x)));
        const float tlb = fmaf(nG, hdx.y, fmaf(nH, hdy.y, fabsf(pv.y)));
        nlo[2 * i] += __float_as_uint(tla) >> 31;
        nlo[2 * i + 1] += __float_as_uint(tlb) >> 31;
        if (UNC == 0) {
          nhi[2 * i] += __float_as_uint(fmaf(nk, fabsf(pv.x), tla)) >> 31;
          nhi[2 * i + 1] += __float_as_uint(fmaf(nk, fabsf(pv.y), tlb)) >> 31;
        } else if (UNC == 1) {
          mn[i] = fminf(fminf(mn[i], fmaf(nk, fabsf(pv.x), fabsf(tla))), fmaf(nk, fabsf(pv.y), fabsf(tlb)));
        } else {
          mn[i] = fminf(fminf(mn[i], fabsf(tla)), fabsf(tlb));
        }
      }
      cx += 0.25f;
      cy -= 0.125f;
    }
  }
  unsigned tot = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += nlo[i] + 3u * nhi[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) tot += __float_as_uint(mn[i]);
  return tot;
}

// The shipped k_score inner loop (chunk-local form): 4 FFMA + FADD + LEA.HI + 1/2 FMNMX3 per unit.
// Exploration knobs: MINMODE 0 none, 1 FMNMX3 per pair (shipped), 2 FMNMX per unit;
//                    CNTMODE 0 none, 1 LEA.HI (shipped), 2 IMAD.HI (FMA pipe), 3 alternate LEA.HI / IMAD.HI
template <int MINMODE, int CNTMODE>
__device__ __forceinline__ unsigned mix_shipped(int iters, float a, float b) {
  float hx[8], hy[8], mn[8];
  unsigned nlo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    hx[i] = a * (float)(i + 1);
    hy[i] = b + (float)i;
    nlo[i] = 0u;
    mn[i] = 3.0e38f;
  }
  float4 A = make_float4(0.6f, -0.8f, a, b);
  float2 B = make_float2(-0.08f, -0.11f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const float p0 = fmaf(A.x, hy[i], fmaf(A.y, hx[i], A.z));
        const float p1 = fmaf(A.x, hy[i + 1], fmaf(A.y, hx[i + 1], A.z));
        const float s0 = fmaf(B.x, hx[i], fmaf(B.y, hy[i], A.w));
        const float s1 = fmaf(B.x, hx[i + 1], fmaf(B.y, hy[i + 1], A.w));
        const float t0v = fabsf(p0) + s0;
        const float t1v = fabsf(p1) + s1;
        if (CNTMODE == 1 || (CNTMODE == 3 && (i & 2))) {
          nlo[i] += __float_as_uint(t0v) >> 31;
          nlo[i + 1] += __float_as_uint(t1v) >> 31;
        } else if (CNTMODE == 2 || CNTMODE == 3) {
          nlo[i] = __umulhi(__float_as_uint(t0v), 2u) + nlo[i];
          nlo[i + 1] = __umulhi(__float_as_uint(t1v), 2u) + nlo[i + 1];
        } else {
          nlo[i] ^= __float_as_uint(t0v + t1v);  // keep t alive with one op per pair
        }
        if (MINMODE == 1) {
          mn[i >> 1] = fminf(fminf(mn[i >> 1], fabsf(t0v)), fabsf(t1v));
        } else if (MINMODE == 2) {
          mn[i] = fminf(mn[i], fabsf(t0v));
          mn[i + 1] = fminf(mn[i + 1], fabsf(t1v));
        }
      }
      A.z += 0.25f;
      A.w -= 0.125f;
    }
  }
  unsigned tot = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += nlo[i] + __float_as_uint(mn[i]);
  return tot;
}

// Packed form of the shipped loop: the four FMAs of two hypotheses as FFMA2 (fma.rn.f32x2).
// PIXU = pixels per iteration (unroll), ADD2 = 1 uses FADD2 for t (then |p| needs a separate abs)
template <int PIXU>
__device__ __forceinline__ unsigned mix_packed(int iters, float a, float b) {
  float2 hx2[4], hy2[4];
  float mn[4];
  unsigned nlo[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    hx2[i] = make_float2(a * (float)(2 * i + 1), a * (float)(2 * i + 2));
    hy2[i] = make_float2(b + (float)(2 * i), b + (float)(2 * i + 1));
    mn[i] = 3.0e38f;
    nlo[2 * i] = nlo[2 * i + 1] = 0u;
  }
  float4 A = make_float4(0.6f, -0.8f, a, b);
  float2 B = make_float2(-0.08f, -0.11f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int q = 0; q < PIXU; ++q) {
      const float2 Ax2 = make_float2(A.x, A.x), Ay2 = make_float2(A.y, A.y), Az2 = make_float2(A.z, A.z);
      const float2 Aw2 = make_float2(A.w, A.w), Bx2 = make_float2(B.x, B.x), By2 = make_float2(B.y, B.y);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 p2 = __ffma2_rn(Ax2, hy2[i], __ffma2_rn(Ay2, hx2[i], Az2));
        const float2 s2 = __ffma2_rn(Bx2, hx2[i], __ffma2_rn(By2, hy2[i], Aw2));
        const float t0v = fabsf(p2.x) + s2.x;
        const float t1v = fabsf(p2.y) + s2.y;
        nlo[2 * i] += __float_as_uint(t0v) >> 31;
        nlo[2 * i + 1] += __float_as_uint(t1v) >> 31;
        mn[i] = fminf(fminf(mn[i], fabsf(t0v)), fabsf(t1v));
      }
      A.z += 0.25f;
      A.w -= 0.125f;
    }
  }
  unsigned tot = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += nlo[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) tot += __float_as_uint(mn[i]);
  return tot;
}

// VARIANT 0: FFMA, 16 independent chains, 2 shared operands
// VARIANT 1: FFMA2 (fma.rn.f32x2), 8 independent float2 chains
// VARIANT 2: FFMA with three distinct register operands per instruction
// VARIANT 3: the scoring loop as shipped; VARIANT 10+PACK*3+UNC: exploration mixes (mix_loop)
template <int VARIANT>
__global__ void __launch_bounds__(256) k_fma_peak(float* out, int iters, float seedv) {
  const float a = 1.0f + seedv * (float)threadIdx.x, b = seedv;
  float r = 0.f;
  if (VARIANT == 0) {
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = (float)k;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 16; ++k) acc[k] = fmaf(acc[k], a, b);
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) r += acc[k];
  } else if (VARIANT == 1) {
    float2 acc[8];
    const float2 a2 = make_float2(a, a + seedv), b2 = make_float2(b, b + seedv);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = make_float2((float)k, (float)k + 0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = __ffma2_rn(acc[k], a2, b2);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) r += acc[k].x + acc[k].y;
  } else if (VARIANT == 2) {
    float acc[16], x[8], y[8];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = (float)k;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      x[k] = a + (float)k * seedv;
      y[k] = b - (float)k * seedv;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 16; ++k) acc[k] = fmaf(x[k & 7], y[(k + 3) & 7], acc[k]);
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) r += acc[k];
  } else if (VARIANT == 3) {
    r = __uint_as_float(mix_shipped<1, 1>(iters, a, b));
  } else if (VARIANT == 40) {
    r = __uint_as_float(mix_packed<2>(iters, a, b));
  } else if (VARIANT >= 20) {
    r = __uint_as_float(mix_shipped<(VARIANT - 20) / 4, (VARIANT - 20) % 4>(iters, a, b));
  } else {
    r = __uint_as_float(mix_loop<(VARIANT - 10) / 3, (VARIANT - 10) % 3>(iters, a, b));
  }
  if (__float_as_uint(r) == 0x7f123456u) out[0] = r;  // keep the work alive
}

}  // namespace casa
