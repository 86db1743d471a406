// tc_epilogue_probe — measures the one number the tensor-core plan of DESIGN.md section 7 rests on: how many
// dispatch cycles per 32 units the EPILOGUE of the scoring loop costs when the two linear forms come out of TMEM
// (tcgen05.ld) instead of four FFMAs per unit.
//
// STATUS: compiles for sm_100a; NOT YET RUN ON HARDWARE (written after the round's GPU minutes were spent).  Not part
// of libcasapose_b200.so, of the tests or of the bench.  Run it AFTER tc_probe (which validates the descriptors this
// file reuses):   bash scripts/tc_probe.sh epilogue     — always under the script's `timeout`.
//
// One CTA per SM, 9 warps.  Warp 8, one elected lane: per iteration ONE tcgen05.mma.kind::tf32 (M 128 hypotheses,
// N 256 = 128 pixels x {p, s}, K 8) into one of two TMEM accumulators (2 x 256 columns), tcgen05.commit -> full[stage].
// Warps 0-7 (two per TMEM lane quarter, each taking 64 of the 128 pixels): wait full[stage], tcgen05.ld p and s,
// per unit  t = |p| + s,  count += sign(t),  min|t| per pair  — the shipped loop's FADD + LEA.HI + 1/2 FMNMX3 —
// then arrive on empty[stage].  The operands never change, so every iteration must produce the same counts: the
// host checks  count == iters x (negative t of that hypothesis over the tile)  against a float64 evaluation
// (entries whose |t| is within 1e-4 of zero are excluded from the check), and reports cycles per 32 units per
// scheduler (4 schedulers per SM; 8 in the FP32 loop today).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

constexpr int kM = 128, kN = 256, kK = 8, kPix = 128;
constexpr uint32_t kLBO = 128, kSBO = 256;
constexpr uint32_t kTmemCols = 512;
constexpr int kEpiWarps = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(kLBO >> 4) << 16;
  d |= (uint64_t)(kSBO >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__host__ __device__ constexpr uint32_t make_idesc() {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
}

__device__ __forceinline__ uint32_t tile_offset(int row, int k) {
  return (uint32_t)((row & 7) * 16 + (row >> 3) * kSBO + (k & 3) * 4 + (k >> 2) * kLBO);
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.b32 %0, 1, 0, q;\n\t}"
                 : "=r"(done)
                 : "r"(bar), "r"(parity)
                 : "memory");
  }
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}

// Sum of the four sign counters.  PRMT form: every negative unit added 0xFFFF to one of the two 16-bit fields of a
// counter, i.e. the counter is -(n_lo) in the low field and -(n_hi) - borrow in the high field.
__device__ __forceinline__ unsigned decode4(const unsigned (&c)[4]) {
  unsigned n = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#ifdef CNT_PRMT
    const unsigned lo = (0u - c[k]) & 0xFFFFu;
    const unsigned hi = (0u - ((c[k] + lo) >> 16)) & 0xFFFFu;  // c + lo has a zero low field (carry into the high one)
    n += lo + hi;
#else
    n += c[k];
#endif
  }
  return n;
}

__global__ void __launch_bounds__((kEpiWarps + 1) * 32, 1)
k_tc_epilogue(const float* __restrict__ A, const float* __restrict__ B, int iters, unsigned* __restrict__ counts,
              float* __restrict__ mins) {
  __shared__ __align__(128) uint8_t sA[kM / 8 * kSBO];
  __shared__ __align__(128) uint8_t sB[kN / 8 * kSBO];
  __shared__ __align__(8) uint64_t full_bar[2], empty_bar[2];
  __shared__ uint32_t tmem_base_smem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int e = tid; e < kM * kK; e += blockDim.x) *(float*)(sA + tile_offset(e / kK, e % kK)) = A[e];
  for (int e = tid; e < kN * kK; e += blockDim.x) *(float*)(sB + tile_offset(e / kK, e % kK)) = B[e];
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full_bar[s])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty_bar[s])), "n"(kEpiWarps));
    }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == kEpiWarps) {
    // ---------------------------------------------------------------- MMA issuer (one lane)
    if (lane == 0) {
      const uint64_t adesc = make_smem_desc(smem_u32(sA));
      const uint64_t bdesc = make_smem_desc(smem_u32(sB));
      const uint32_t idesc = make_idesc();
      for (int it = 0; it < iters; ++it) {
        const int s = it & 1;
        mbar_wait(smem_u32(&empty_bar[s]), (uint32_t)(((it >> 1) & 1) ^ 1));  // first two waits pass immediately
        asm volatile("tcgen05.fence::after_thread_sync;");
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_base + (uint32_t)(s * kN)), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(0u));
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&full_bar[s])));
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue: thread = one hypothesis, 64 pixels
    const int quarter = warp & 3, half = warp >> 2;
    unsigned cnt4[4] = {0u, 0u, 0u, 0u};  // four independent chains: one thread has no other ILP here
    unsigned total = 0u;
    float mn4[4] = {3.0e38f, 3.0e38f, 3.0e38f, 3.0e38f};
    for (int it = 0; it < iters; ++it) {
      const int s = it & 1;
      mbar_wait(smem_u32(&full_bar[s]), (uint32_t)((it >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;");
      const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(s * kN);
#pragma unroll
      for (int j0 = 0; j0 < kPix / 2; j0 += 32) {
        uint32_t p[32], q[32];
        tmem_ld32(lane_base + (uint32_t)(half * 64 + j0), p);          // p of 32 pixels
        tmem_ld32(lane_base + (uint32_t)(kPix + half * 64 + j0), q);   // s of the same pixels
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#ifdef CNT_PRMT
        // sign count with PRMT (sign-replicating byte permute: 0xFFFF per negative t, two units) + IADD3 over two
        // pairs: 3 ALU-pipe instructions per 4 units instead of 4 LEA.HI; the two 16-bit fields are decoded at the end
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float t[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) t[u] = fabsf(__uint_as_float(p[j + u])) + __uint_as_float(q[j + u]);
          unsigned r0, r1;
          asm("prmt.b32 %0, %1, %2, 0xFFBB;" : "=r"(r0) : "r"(__float_as_uint(t[0])), "r"(__float_as_uint(t[1])));
          asm("prmt.b32 %0, %1, %2, 0xFFBB;" : "=r"(r1) : "r"(__float_as_uint(t[2])), "r"(__float_as_uint(t[3])));
          cnt4[(j >> 2) & 3] = cnt4[(j >> 2) & 3] + r0 + r1;
          mn4[(j >> 2) & 3] = fminf(fminf(mn4[(j >> 2) & 3], fabsf(t[0])), fabsf(t[1]));
          mn4[((j >> 2) + 2) & 3] = fminf(fminf(mn4[((j >> 2) + 2) & 3], fabsf(t[2])), fabsf(t[3]));
        }
#else
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float t0 = fabsf(__uint_as_float(p[j])) + __uint_as_float(q[j]);
          const float t1 = fabsf(__uint_as_float(p[j + 1])) + __uint_as_float(q[j + 1]);
          cnt4[(j >> 1) & 3] += __float_as_uint(t0) >> 31;
          cnt4[((j >> 1) + 2) & 3] += __float_as_uint(t1) >> 31;
          mn4[(j >> 1) & 3] = fminf(fminf(mn4[(j >> 1) & 3], fabsf(t0)), fabsf(t1));
        }
#endif
      }
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty_bar[s])) : "memory");
      if ((it & 255) == 255) {  // the 16-bit fields of the PRMT form hold 256 tiles of 64 pixels
        total += decode4(cnt4);
        cnt4[0] = cnt4[1] = cnt4[2] = cnt4[3] = 0u;
      }
    }
    const unsigned cnt = total + decode4(cnt4);
    const float mn = fminf(fminf(mn4[0], mn4[1]), fminf(mn4[2], mn4[3]));
    const int hyp = quarter * 32 + lane;
    counts[((size_t)blockIdx.x * kM + hyp) * 2 + half] = cnt;
    mins[((size_t)blockIdx.x * kM + hyp) * 2 + half] = mn;
  }

  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
}

static float tf32_round(float x) {
  uint32_t b;
  memcpy(&b, &x, 4);
  b = (b + 0x1000u) & 0xFFFFE000u;
  float y;
  memcpy(&y, &b, 4);
  return y;
}

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 4000;
  static float hA[kM * kK], hB[kN * kK];
  srand(2);
  for (int m = 0; m < kM; ++m) {
    const float hx = 200.f * ((float)rand() / RAND_MAX - 0.5f), hy = 200.f * ((float)rand() / RAND_MAX - 0.5f);
    const float xh = tf32_round(hx), xl = tf32_round(hx - xh), yh = tf32_round(hy), yl = tf32_round(hy - yh);
    const float row[kK] = {xh, xl, xh, yh, yl, yh, 1.f, 1.f};
    memcpy(&hA[m * kK], row, sizeof(row));
  }
  const float k_lo = 0.1424f;
  for (int n = 0; n < kPix; ++n) {  // a pixel at (cx, cy) with unit direction (D, E): p row n, s row kPix + n
    const float ang = 6.2831853f * (float)rand() / RAND_MAX, D = cosf(ang), E = sinf(ang);
    const float cx = 60.f * ((float)rand() / RAND_MAX - 0.5f), cy = 4.f * ((float)rand() / RAND_MAX - 0.5f);
    const float G = k_lo * D, H = k_lo * E, P0 = D * cy - E * cx, A0 = G * cx + H * cy;
    const float v[6] = {-E, D, -P0, -G, -H, A0};
    float hi[6], lo[6];
    for (int i = 0; i < 6; ++i) {
      hi[i] = tf32_round(v[i]);
      lo[i] = tf32_round(v[i] - hi[i]);
    }
    const float prow[kK] = {hi[0], hi[0], lo[0], hi[1], hi[1], lo[1], hi[2], lo[2]};
    const float srow[kK] = {hi[3], hi[3], lo[3], hi[4], hi[4], lo[4], hi[5], lo[5]};
    memcpy(&hB[n * kK], prow, sizeof(prow));
    memcpy(&hB[(kPix + n) * kK], srow, sizeof(srow));
  }
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int clock_khz = 0;
  CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, dev));
  float *dA, *dB, *dMin;
  unsigned* dCnt;
  CK(cudaMalloc(&dA, sizeof(hA)));
  CK(cudaMalloc(&dB, sizeof(hB)));
  CK(cudaMalloc(&dCnt, (size_t)sms * kM * 2 * sizeof(unsigned)));
  CK(cudaMalloc(&dMin, (size_t)sms * kM * 2 * sizeof(float)));
  CK(cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  k_tc_epilogue<<<sms, (kEpiWarps + 1) * 32>>>(dA, dB, 16, dCnt, dMin);  // warm-up
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  k_tc_epilogue<<<sms, (kEpiWarps + 1) * 32>>>(dA, dB, iters, dCnt, dMin);
  CK(cudaEventRecord(e1));
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  unsigned* hC = (unsigned*)malloc((size_t)sms * kM * 2 * sizeof(unsigned));
  CK(cudaMemcpy(hC, dCnt, (size_t)sms * kM * 2 * sizeof(unsigned), cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int m = 0; m < kM; ++m)
    for (int half = 0; half < 2; ++half) {
      int neg = 0, fuzzy = 0;
      for (int n = half * 64; n < half * 64 + 64; ++n) {
        double p = 0.0, s = 0.0;
        for (int k = 0; k < kK; ++k) {
          p += (double)hA[m * kK + k] * (double)hB[n * kK + k];
          s += (double)hA[m * kK + k] * (double)hB[(kPix + n) * kK + k];
        }
        const double t = fabs(p) + s;
        if (fabs(t) < 1e-4) ++fuzzy;
        if (t < 0) ++neg;
      }
      for (int b = 0; b < sms; ++b) {
        const long long got = hC[((size_t)b * kM + m) * 2 + half], want = (long long)neg * iters;
        if (llabs(got - want) > (long long)fuzzy * iters) ++bad;
      }
    }
  const double units = (double)sms * iters * kM * kPix;
  const double cycles = (double)ms * 1e-3 * (double)clock_khz * 1e3;
  printf("tc_epilogue_probe: %d SMs x %d iterations, %.3f ms, %.1f G units/s, %.2f cycles per 32 units per scheduler "
         "(FP32 loop: 8), wrong counters: %d\n",
         sms, iters, ms, units / ms * 1e-6, cycles * 4.0 / ((double)iters * kM * kPix / 32.0), bad);
  if (bad == 0) printf("tc_epilogue_probe ok\n");
  return bad != 0;
}
