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
    const bool ex = exact_inlier(hx, hy, cx, cy, dx, dy, exact_norm(dx, dy), fc.thr);
    ++tested;
    inl += ex;
    // ---- the chunk-local form exactly as k_score evaluates it: random chunk origin within 90 px of the pixel
    // (one fractional bit, like a bounding-box centre), chunk radius >= this pixel's offset
    const float ox = cx + 0.5f * (float)((int)(r1.w % 361u) - 180), oy = cy + 0.5f * (float)((int)((r1.w >> 10) % 361u) - 180);
    const float cxl = cx - ox, cyl = cy - oy;
    float4 A;
    float2 B;
    if (!make_local_coef(cxl, cyl, dx, dy, fc.k_mid, A, B)) continue;
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

// FP32 issue-rate micro-benchmark: FFMA, 16 independent chains per thread, 2 shared operands.  Its rate is the
// denominator of "% of FP32 peak" (MEASURED_PEAKS.json has no FP32 figure).  The instruction-mix explorations that
// used to live here are in csrc/experimental/ (fma_mix_round1.cuh, loop_bench.cu) — not part of the library.
__global__ void __launch_bounds__(256) k_fma_peak(float* out, int iters, float seedv) {
  const float a = 1.0f + seedv * (float)threadIdx.x, b = seedv;
  float r = 0.f;
  float acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = (float)k;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = fmaf(acc[k], a, b);
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) r += acc[k];
  if (__float_as_uint(r) == 0x7f123456u) out[0] = r;  // keep the work alive
}

}  // namespace casa
