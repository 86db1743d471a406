// tc_probe — a stand-alone probe for the tensor-core form of the scoring loop.
//
// STATUS: compiles for sm_100a (ptxas accepts every tcgen05 form below); NOT YET RUN ON HARDWARE — the GPU budget of
// round 1 was spent before it was written.  It is not part of libcasapose_b200.so, of the tests or of the bench.
// First GPU step of the next round:   bash scripts/tc_probe.sh   (runs it under `timeout 20`).
//
// What it checks: ONE tcgen05.mma.cta_group::1.kind::tf32 with M = 128 (hypotheses), N = 256 (128 pixels x the two
// linear forms p | s), K = 8 (the 3xTF32 split of DESIGN.md), operands written to shared memory by ordinary stores in
// the no-swizzle K-major core-matrix order, accumulator in TMEM, read back with tcgen05.ld.32x32b — against a float64
// host evaluation of the same eight products.  If the descriptors below are right the program prints the maximum
// deviation (expected: a few float32 ulps of the row scale) and "tc_probe ok".
//
// Shared-memory operand layout (no swizzle, K-major; element = 4-byte TF32 container):
//   a core matrix is 8 rows x 16 bytes (4 elements), stored as 128 contiguous bytes (row r at r*16);
//   element (row, k) of a tile lives at  (row % 8) * 16 + (row / 8) * SBO + (k % 4) * 4 + (k / 4) * LBO
//   with LBO = 128 (the second K core matrix follows the first) and SBO = 256 (next group of 8 rows).
// Matrix descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp): start address >> 4 in bits [0,14), LBO >> 4 in
// [16,30), SBO >> 4 in [32,46), version = 1 in [46,48), layout type SWIZZLE_NONE = 0 in [61,64).
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 at [4,6), a_format = b_format = TF32 = 2 at
// [7,10) and [10,13), a_major = b_major = K (0), n_dim = N >> 3 at [17,23), m_dim = M >> 4 at [24,29).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

constexpr int kM = 128, kN = 256, kK = 8;
constexpr uint32_t kLBO = 128, kSBO = 256;
constexpr uint32_t kTmemCols = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version of sm_100
  return d;                // base offset 0, LBO mode 0, SWIZZLE_NONE
}

__host__ __device__ constexpr uint32_t make_idesc() {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
}

__device__ __forceinline__ uint32_t tile_offset(int row, int k) {  // byte offset inside an operand tile
  return (uint32_t)((row & 7) * 16 + (row >> 3) * kSBO + (k & 3) * 4 + (k >> 2) * kLBO);
}

__global__ void __launch_bounds__(128, 1) k_tc_probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, uint32_t lbo, uint32_t sbo) {
  __shared__ __align__(128) uint8_t sA[kM / 8 * kSBO];  // 4 KB
  __shared__ __align__(128) uint8_t sB[kN / 8 * kSBO];  // 8 KB
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  const int tid = threadIdx.x, warp = tid >> 5;

  // operands: A[kM][kK], B[kN][kK] row-major in global memory -> core-matrix order in shared memory
  for (int k = 0; k < kK; ++k) {
    *(float*)(sA + tile_offset(tid, k)) = A[tid * kK + k];
    *(float*)(sB + tile_offset(tid, k)) = B[tid * kK + k];
    *(float*)(sB + tile_offset(tid + 128, k)) = B[(tid + 128) * kK + k];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {  // one warp allocates the accumulator columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");  // generic-proxy stores above -> visible to the tensor core's async proxy
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = tmem_base_smem;

  if (tid == 0) {  // a single thread issues the MMA on behalf of the CTA
    const uint64_t adesc = make_smem_desc(smem_u32(sA), lbo, sbo);
    const uint64_t bdesc = make_smem_desc(smem_u32(sB), lbo, sbo);
    const uint32_t idesc = make_idesc();
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_base), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(0u));  // p = 0: D = A * B^T (no accumulate)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
  }
  // everybody waits for the accumulator (phase 0 of the barrier)
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.b32 %0, 1, 0, q;\n\t}"
          : "=r"(done)
          : "r"(smem_u32(&bar)), "r"(0u));
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;");

  // TMEM lane = row of D (hypothesis), column = column of D; warp w reads lanes 32w .. 32w+31, 32 columns at a time
  const int row = tid;
#pragma unroll 1
  for (int c0 = 0; c0 < kN; c0 += 32) {
    uint32_t v[32];
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
    for (int j = 0; j < 32; ++j) D[row * kN + c0 + j] = __uint_as_float(v[j]);
  }

  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
}

static float tf32_round(float x) {  // round to nearest, ties away (cvt.rna.tf32.f32)
  uint32_t b;
  memcpy(&b, &x, 4);
  b = (b + 0x1000u) & 0xFFFFE000u;
  float y;
  memcpy(&y, &b, 4);
  return y;
}

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 1;                                                                    \
    }                                                                              \
  } while (0)

int main(int argc, char** argv) {
  const bool swap = argc > 1 && !strcmp(argv[1], "swap");  // descriptor fields exchanged (layout in memory unchanged)
  static float hA[kM * kK], hB[kN * kK], hD[kM * kN];
  srand(1);
  // A rows: [hx_hi, hx_lo, hx_hi, hy_hi, hy_lo, hy_hi, 1, 1]; B rows: p | s coefficient splits (random stand-ins)
  for (int m = 0; m < kM; ++m) {
    const float hx = 200.f * ((float)rand() / RAND_MAX - 0.5f), hy = 200.f * ((float)rand() / RAND_MAX - 0.5f);
    const float xh = tf32_round(hx), xl = tf32_round(hx - xh), yh = tf32_round(hy), yl = tf32_round(hy - yh);
    const float row[kK] = {xh, xl, xh, yh, yl, yh, 1.f, 1.f};
    for (int k = 0; k < kK; ++k) hA[m * kK + k] = row[k];
  }
  for (int n = 0; n < kN; ++n) {
    const float a = (float)rand() / RAND_MAX - 0.5f, b = (float)rand() / RAND_MAX - 0.5f, c = 50.f * ((float)rand() / RAND_MAX - 0.5f);
    const float ah = tf32_round(a), al = tf32_round(a - ah), bh = tf32_round(b), bl = tf32_round(b - bh);
    const float ch = tf32_round(c), cl = tf32_round(c - ch);
    const float row[kK] = {ah, ah, al, bh, bh, bl, ch, cl};
    for (int k = 0; k < kK; ++k) hB[n * kK + k] = row[k];
  }
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, sizeof(hA)));
  CK(cudaMalloc(&dB, sizeof(hB)));
  CK(cudaMalloc(&dD, sizeof(hD)));
  CK(cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, sizeof(hD)));
  k_tc_probe<<<1, 128>>>(dA, dB, dD, swap ? kSBO : kLBO, swap ? kLBO : kSBO);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost));
  double worst = 0.0;
  int bad = 0;
  for (int m = 0; m < kM; ++m)
    for (int n = 0; n < kN; ++n) {
      double ref = 0.0, scale = 0.0;
      for (int k = 0; k < kK; ++k) {
        ref += (double)hA[m * kK + k] * (double)hB[n * kK + k];
        scale += fabs((double)hA[m * kK + k] * (double)hB[n * kK + k]);
      }
      const double err = fabs((double)hD[m * kN + n] - ref) / (scale + 1e-30);
      if (!(err < 1e-5)) ++bad;
      if (err > worst || err != err) worst = err;
    }
  printf("tc_probe%s: max |D - ref| / sum|terms| = %.3g (%.2f u), %d of %d entries off\n", swap ? " (LBO/SBO fields swapped)" : "", worst, worst / 5.96e-8, bad, kM * kN);
  if (bad == 0) printf("tc_probe ok\n");
  return bad != 0;
}
