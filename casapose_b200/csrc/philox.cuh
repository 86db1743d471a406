// Philox4x32-10 (Salmon et al., SC'11) and the two random streams of the voting path.
// Bit-identical restatement of oracle/philox_np.py; stands in for tf.random.uniform at
// /root/reference/casapose/pose_estimation/ransac_voting.py:296 (selection) and :319 (idxs).
#pragma once
#include "common.cuh"

namespace casa {

constexpr uint32_t kStreamIdxs = 0;
constexpr uint32_t kStreamSelection = 1;

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

__device__ __forceinline__ uint32_t pick(uint4 r, uint32_t lane) {
  return lane == 0 ? r.x : lane == 1 ? r.y : lane == 2 ? r.z : r.w;
}

// word e of stream (c1, c2, c3)
__device__ __forceinline__ uint32_t philox_word(uint32_t e, uint32_t c1, uint32_t c2, uint32_t c3,
                                                uint32_t k0, uint32_t k1) {
  return pick(philox4x32_10(make_uint4(e >> 2, c1, c2, c3), k0, k1), e & 3u);
}

// idx pair of hypothesis (h, v) in round r of job (image, cls):  idx = (word * tn) >> 32
__device__ __forceinline__ int2 philox_idx_pair(uint32_t h, uint32_t v, uint32_t vn, uint32_t rnd,
                                                uint32_t cls, uint32_t image, uint32_t tn,
                                                uint32_t k0, uint32_t k1) {
  const uint32_t e = (h * vn + v) * 2u;  // even -> both words live in the same Philox block
  const uint4 r = philox4x32_10(make_uint4(e >> 2, rnd, (cls & 0xFFFFu) | (kStreamIdxs << 16), image), k0, k1);
  const uint32_t w0 = (e & 2u) ? r.z : r.x;
  const uint32_t w1 = (e & 2u) ? r.w : r.y;
  return make_int2((int)__umulhi(w0, tn), (int)__umulhi(w1, tn));
}

// selection value of pixel e = y*w + x of job (image, cls): uniform float32 in [0,1)
__device__ __forceinline__ float philox_selection(uint32_t e, uint32_t cls, uint32_t image,
                                                  uint32_t k0, uint32_t k1) {
  const uint32_t wd = philox_word(e, 0u, (cls & 0xFFFFu) | (kStreamSelection << 16), image, k0, k1);
  return (float)(wd >> 8) * 5.9604644775390625e-8f;  // * 2^-24, exact
}

}  // namespace casa
