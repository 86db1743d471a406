// Host-side mask packer in isolation: bytes/s of the AVX2 loop of casa_api.cu (MaskPacker::pack8_avx2) on pageable memory,
// 1..16 threads, plus a plain read (sum) of the same buffer as the memory-bandwidth reference.
// build: g++ -O3 -mavx2 -pthread -o pack_bench pack_bench.cpp
#include <immintrin.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <thread>
#include <vector>

static unsigned pack8(const float* mask, uint32_t* bits, size_t lo, size_t hi) {
  const __m256i zero = _mm256_setzero_si256(), one = _mm256_set1_epi32(0x3F800000);
  unsigned bad = 0;
  for (size_t p = lo; p < hi; ++p) {
    const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(mask + 8 * p));
    const unsigned z = (unsigned)_mm256_movemask_ps(_mm256_castsi256_ps(_mm256_cmpeq_epi32(_mm256_slli_epi32(v, 1), zero)));
    const unsigned e1 = (unsigned)_mm256_movemask_ps(_mm256_castsi256_ps(_mm256_cmpeq_epi32(v, one)));
    const unsigned m = ~z & 0xFFu;
    bad |= m & ~e1;
    bits[p] = m;
  }
  return bad;
}
// 8 pixels per trip: one 256-bit OR-reduction decides whether the 8 rows are all zero (background: 87 % of the rows)
static unsigned pack8_skip(const float* mask, uint32_t* bits, size_t lo, size_t hi) {
  const __m256i zero = _mm256_setzero_si256(), one = _mm256_set1_epi32(0x3F800000);
  unsigned bad = 0;
  size_t p = lo;
  for (; p + 8 <= hi; p += 8) {
    const __m256i* r = reinterpret_cast<const __m256i*>(mask + 8 * p);
    __m256i v[8];
    for (int k = 0; k < 8; ++k) v[k] = _mm256_loadu_si256(r + k);
    const __m256i any = _mm256_or_si256(_mm256_or_si256(_mm256_or_si256(v[0], v[1]), _mm256_or_si256(v[2], v[3])),
                                        _mm256_or_si256(_mm256_or_si256(v[4], v[5]), _mm256_or_si256(v[6], v[7])));
    if (_mm256_testz_si256(any, any)) {
      _mm256_storeu_si256(reinterpret_cast<__m256i*>(bits + p), zero);
      continue;
    }
    for (int k = 0; k < 8; ++k) {
      const unsigned z = (unsigned)_mm256_movemask_ps(_mm256_castsi256_ps(_mm256_cmpeq_epi32(_mm256_slli_epi32(v[k], 1), zero)));
      const unsigned e1 = (unsigned)_mm256_movemask_ps(_mm256_castsi256_ps(_mm256_cmpeq_epi32(v[k], one)));
      const unsigned m = ~z & 0xFFu;
      bad |= m & ~e1;
      bits[p + k] = m;
    }
  }
  if (p < hi) bad |= pack8(mask, bits, p, hi);
  return bad;
}
static unsigned long long readsum(const float* mask, size_t lo, size_t hi) {
  __m256i acc = _mm256_setzero_si256();
  for (size_t p = lo; p < hi; ++p) acc = _mm256_or_si256(acc, _mm256_loadu_si256(reinterpret_cast<const __m256i*>(mask + 8 * p)));
  unsigned long long o[4];
  _mm256_storeu_si256(reinterpret_cast<__m256i*>(o), acc);
  return o[0] | o[1] | o[2] | o[3];
}
int main() {
  const size_t npx = (size_t)16 * 480 * 640;
  float* mask = (float*)aligned_alloc(64, npx * 32);
  uint32_t* bits = (uint32_t*)aligned_alloc(64, npx * 4);
  memset(mask, 0, npx * 32);
  for (size_t p = 0; p < npx; ++p)
    if ((p / 97) % 8 == 0) mask[8 * p + (p % 8)] = 1.0f;  // 12.5 % masked, in runs
  memset(bits, 0, npx * 4);
  printf("hardware_concurrency %u\n", std::thread::hardware_concurrency());
  for (int mode = 0; mode < 3; ++mode)
    for (int n : {1, 2, 4, 8, 12, 16}) {
      double best = 1e9;
      for (int rep = 0; rep < 5; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        volatile unsigned long long sink = 0;
        for (int i = 0; i < n; ++i)
          th.emplace_back([&, i] {
            const size_t lo = npx * i / n, hi = npx * (i + 1) / n;
            unsigned long long r = mode == 0 ? readsum(mask, lo, hi) : (mode == 1 ? pack8(mask, bits, lo, hi) : pack8_skip(mask, bits, lo, hi));
            sink = sink | r;
          });
        for (auto& t : th) t.join();
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        best = s < best ? s : best;
      }
      printf("%s threads %2d: %.2f ms  %.1f GB/s\n", mode == 0 ? "read " : (mode == 1 ? "pack " : "pack8"), n, best * 1e3, npx * 32 / best / 1e9);
    }
  return 0;
}
