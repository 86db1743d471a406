// loop_bench — stand-alone instruction-mix benchmark for the inner loop of k_score (ransac.cuh).
//
// Not part of libcasapose_b200.so, of the tests or of the bench: design exploration only
// (`bash scripts/loop_bench.sh` on a B200).  Every variant evaluates the same units as the shipped loop
//     p = D hy' - E hx' - P0,  s = A0 - G hx' - H hy',  t = |p| + s,  count += sign(t),  min|t| per pair
// with the pixel coefficients broadcast from shared memory (LDS.128 + LDS.64 per pixel) and 8 hypotheses per lane;
// what changes is HOW the sign is counted and the minimum tracked:
//   0  FFMA only (16 independent chains): the FP32 issue peak of this box
//   1  shipped: LEA.HI per unit + FMNMX3 per pair
//   2  PRMT (sign-replicating byte permute, 2 units) + IADD3 over two pixels (4 units) + FMNMX3 per pair
//   3  shipped without the minimum          4  shipped without the count
//   5  count on the FP32 pipe: FMUL.SAT (indicator) + FADD
//   6  PRMT + integer multiply-add accumulate (IMAD runs on the FMA pipe)
//   7  no count, no minimum (4 FFMA + FADD + one XOR per pair keeps t alive): the floor of the formulation
//   8  as 2 with 16 hypotheses per lane
//   9  IMAD.HI.U32 count (hi32(t * 2) + n, FMA pipe) + FMNMX3 per pair
//  10  as 6 with the trip unrolled to 4 pixels (loop overhead halved)
//  11  IMAD.HI count, no minimum        12  IMAD.WIDE count (64-bit accumulator; wrong counts, rate probe only)
// Prints TFLOP/s-equivalent (11 FLOP per unit, the accounting of SURVEY.md section 8d) and cycles per 32 units.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int kPix = 128;

__device__ __forceinline__ unsigned prmt_sign2(float a, float b) {  // 0xFFFF in the low / high half if a / b is negative
  unsigned r;
  asm("prmt.b32 %0, %1, %2, 0xFFBB;" : "=r"(r) : "r"(__float_as_uint(a)), "r"(__float_as_uint(b)));
  return r;
}

template <int V, int HPL>
__global__ void __launch_bounds__(256) k_loop(float* out, int iters, float seedv, unsigned one, unsigned two) {
  __shared__ float4 sA[8][kPix];
  __shared__ float2 sB[8][kPix];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int q = lane; q < kPix; q += 32) {
    sA[warp][q] = make_float4(0.6f + seedv * q, -0.8f, 3.f - 0.01f * q, 1.f + 0.02f * q);
    sB[warp][q] = make_float2(-0.08f, -0.11f + seedv * q);
  }
  __syncwarp();
  const float4* cA = sA[warp];
  const float2* cB = sB[warp];
  if (V == 0) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = seedv * (float)(i + lane);
    const float a = 1.0f + seedv, b = seedv;
    for (int it = 0; it < iters * kPix / 2; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) tot += acc[i];
    if (tot == 123.456f) out[0] = tot;
    return;
  }
  float hx[HPL], hy[HPL], mn[HPL / 2], fc[HPL];
  unsigned nlo[HPL];
#pragma unroll
  for (int i = 0; i < HPL; ++i) {
    hx[i] = seedv * (float)(i + 1) + (float)lane;
    hy[i] = 2.f + (float)i - 0.5f * (float)lane;
    nlo[i] = 0u;
    fc[i] = 0.f;
  }
#pragma unroll
  for (int i = 0; i < HPL / 2; ++i) mn[i] = 3.0e38f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll 1
    for (int q = 0; q < kPix; q += 2) {
      const float4 A0 = cA[q], A1 = cA[q + 1];
      const float2 B0 = cB[q], B1 = cB[q + 1];
#pragma unroll
      for (int i = 0; i < HPL; i += 2) {
        float t[2][2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float4 A = u ? A1 : A0;
          const float2 B = u ? B1 : B0;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const float p = fmaf(A.x, hy[i + j], fmaf(A.y, hx[i + j], A.z));
            t[u][j] = fabsf(p) + fmaf(B.x, hx[i + j], fmaf(B.y, hy[i + j], A.w));
          }
        }
        if (V == 1 || V == 3) {
#pragma unroll
          for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int j = 0; j < 2; ++j) nlo[i + j] += __float_as_uint(t[u][j]) >> 31;
        } else if (V == 2 || V == 8) {
          nlo[i] = nlo[i] + prmt_sign2(t[0][0], t[0][1]) + prmt_sign2(t[1][0], t[1][1]);
        } else if (V == 5) {
#pragma unroll
          for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int j = 0; j < 2; ++j) fc[i + j] += __saturatef(t[u][j] * -1.0e30f);
        } else if (V == 6) {
          nlo[i] = prmt_sign2(t[0][0], t[0][1]) * one + nlo[i];
          nlo[i + 1] = prmt_sign2(t[1][0], t[1][1]) * one + nlo[i + 1];
        } else if (V == 9 || V == 11) {
#pragma unroll
          for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int j = 0; j < 2; ++j) nlo[i + j] = __umulhi(__float_as_uint(t[u][j]), two) + nlo[i + j];
        } else if (V == 7) {
          nlo[i] ^= __float_as_uint(t[0][0] + t[0][1]) ^ __float_as_uint(t[1][0] + t[1][1]);
        }
        if (V != 3 && V != 7 && V != 11) {
          mn[i >> 1] = fminf(fminf(mn[i >> 1], fabsf(t[0][0])), fabsf(t[0][1]));
          mn[i >> 1] = fminf(fminf(mn[i >> 1], fabsf(t[1][0])), fabsf(t[1][1]));
        }
      }
    }
  }
  unsigned tot = 0;
#pragma unroll
  for (int i = 0; i < HPL; ++i) tot += nlo[i] + __float_as_uint(fc[i]);
#pragma unroll
  for (int i = 0; i < HPL / 2; ++i) tot += __float_as_uint(mn[i]);
  if (tot == 0x12345678u) out[0] = (float)tot;
}

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

template <int V, int HPL>
static int run(const char* name, int sms, int blocks_per_sm, float* dout, double peak_tflops, double* tflops_out) {
  const int iters = V == 0 ? 64 : 64 * 8 / HPL, blocks = sms * blocks_per_sm;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    CK(cudaEventRecord(e0));
    k_loop<V, HPL><<<blocks, 256>>>(dout, iters, 1e-9f, 1u, 2u);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  const double thread_iters = (double)blocks * 256.0 * iters * kPix;
  // V == 0: 8 FMA per thread per pixel-step; otherwise HPL units of 11 FLOP per thread per pixel
  const double flop = V == 0 ? thread_iters * 8.0 * 2.0 : thread_iters * HPL * 11.0;
  const double tf = flop / (best * 1e-3) / 1e12;
  if (tflops_out) *tflops_out = tf;
  if (V == 0) {
    printf("%-44s %8.3f ms  %6.2f TFLOP/s\n", name, best, tf);
  } else {
    // cycles per 32 units per scheduler, in units of the measured FFMA issue rate (1 FFMA = 1 cycle)
    const double cyc = 11.0 / 2.0 * peak_tflops / tf;
    printf("%-44s %8.3f ms  %6.2f TFLOP/s-eq  %.3f of FFMA peak  %.2f FFMA-cycles per 32 units\n", name, best, tf,
           tf / peak_tflops, cyc);
  }
  return 0;
}

int main(int argc, char** argv) {
  const int bps = argc > 1 ? atoi(argv[1]) : 3;  // resident blocks per SM (the shipped kernel runs 3 x 256 threads)
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  float* dout;
  CK(cudaMalloc(&dout, 4));
  double peak = 0;
  if (run<0, 8>("0 FFMA peak", sms, bps, dout, 0, &peak)) return 1;
  run<1, 8>("1 shipped (LEA.HI + FMNMX3/2)", sms, bps, dout, peak, nullptr);
  run<2, 8>("2 PRMT.sign + IADD3 + FMNMX3/2", sms, bps, dout, peak, nullptr);
  run<3, 8>("3 LEA.HI, no min", sms, bps, dout, peak, nullptr);
  run<4, 8>("4 FMNMX3/2, no count", sms, bps, dout, peak, nullptr);
  run<5, 8>("5 FMUL.SAT + FADD count + FMNMX3/2", sms, bps, dout, peak, nullptr);
  run<6, 8>("6 PRMT.sign + IMAD + FMNMX3/2", sms, bps, dout, peak, nullptr);
  run<7, 8>("7 no count, no min (floor)", sms, bps, dout, peak, nullptr);
  run<8, 16>("8 as 2, 16 hypotheses per lane", sms, bps, dout, peak, nullptr);
  run<9, 8>("9 IMAD.HI count + FMNMX3/2", sms, bps, dout, peak, nullptr);
  run<11, 8>("11 IMAD.HI count, no min", sms, bps, dout, peak, nullptr);
  return 0;
}
