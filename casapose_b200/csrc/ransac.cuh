// K2-K4 — the RANSAC loop of ransac_voting_batch
// (/root/reference/casapose/pose_estimation/ransac_voting.py:310-368) for every (image, class) job.
//
//   k_hypgen   idxs -> hypotheses, exact float32 sequence (:319-322, :197-227), classifies
//              each hypothesis for the filtered predicate and zeroes the vote counters;
//   k_plan     turns the active jobs into a flat list of scoring work items
//              (job, keypoint, pixel tile) for the persistent scoring grid;
//   k_score    THE HOT KERNEL: hypotheses x pixels inlier test (:230-249) and vote counts (:327);
//   k_update   arg-max per keypoint (:328-333), best-so-far update (:336-338), stop test (:340-347);
//   k_refine   re-vote of the winners and the normal-equation sums (:349-362);
//   k_solve    invertibility test and 2x2 solve (:254-272, :364-368).
#pragma once
#include "common.cuh"
#include "philox.cuh"
#include "predicate.cuh"

namespace casa {

__device__ __forceinline__ float2 load_dir(const float* __restrict__ vimg, int w, int vn, int x, int y, int v) {
  // vertex[img, y, x, v, 0:2] = (dy, dx)  ->  (dx, dy) like tf.reverse(..., axis=[2]) at :308
  const float2 t = __ldg(reinterpret_cast<const float2*>(vimg + ((size_t)(y * w + x) * vn + v) * 2));
  return make_float2(t.y, t.x);
}

// ------------------------------------------------------------------------------------ K2
// one block per job
__global__ void __launch_bounds__(256) k_hypgen(WS ws, Dims d, FilterConsts fc, const float* __restrict__ vertex,
                                                const int32_t* __restrict__ idxs, int rnd, float* dbg_hyps) {
  const int job = blockIdx.x, tid = threadIdx.x;
  int flags = ws.job_flags[job];
  const int tn = ws.job_tn[job];
  if (rnd == 0) {
    const bool act = !(flags & JOB_GATED) && tn > 0;
    __syncthreads();  // everyone has read the old flags
    if (tid == 0) ws.job_flags[job] = act ? (flags | JOB_ACTIVE) : (flags & ~JOB_ACTIVE);
    for (int v = tid; v < d.vn; v += 256) {
      ws.win_ratio[job * d.vn + v] = 0.f;  // :311-312
      ws.win_pts[job * d.vn + v] = make_float2(0.f, 0.f);
    }
    if (!act) return;
  } else if (!(flags & JOB_ACTIVE)) {
    return;
  }
  const int img = job / d.oc, cls = job - img * d.oc;
  const uint32_t* pix = ws.pix + (size_t)img * d.cap + ws.job_off[job];
  const float* vimg = vertex + (size_t)img * d.hw * d.vn * 2;
  for (int v = tid; v < d.vn; v += 256) ws.n_exact[job * d.vn + v] = 0;
  __syncthreads();
  const int n = d.hn * d.vn;
  for (int e = tid; e < n; e += 256) {
    const int h = e / d.vn, v = e - h * d.vn;
    int2 ip;
    if (idxs) {
      const int32_t* src = idxs + ((((size_t)job * d.max_iter + rnd) * d.hn + h) * d.vn + v) * 2;
      ip = make_int2(src[0], src[1]);
      if ((unsigned)ip.x >= (unsigned)tn || (unsigned)ip.y >= (unsigned)tn) {
        atomicOr(reinterpret_cast<unsigned*>(&ws.ctrl[CTRL_STATUS]), CASA_STATUS_IDX_RANGE);
        ip.x = min(max(ip.x, 0), tn - 1);
        ip.y = min(max(ip.y, 0), tn - 1);
      }
    } else {
      ip = philox_idx_pair((uint32_t)h, (uint32_t)v, (uint32_t)d.vn, (uint32_t)rnd, (uint32_t)cls,
                           (uint32_t)(d.image_offset + img), (uint32_t)tn, d.seed_lo, d.seed_hi);
    }
    const uint32_t p0 = pix[ip.x], p1 = pix[ip.y];
    const int x0 = p0 & 0xFFFFu, y0 = p0 >> 16, x1 = p1 & 0xFFFFu, y1 = p1 >> 16;
    const float2 c0 = make_float2((float)x0 + 0.5f, (float)y0 + 0.5f);  // :306
    const float2 c1 = make_float2((float)x1 + 0.5f, (float)y1 + 0.5f);
    const float2 d0 = load_dir(vimg, d.w, d.vn, x0, y0, v);
    const float2 d1 = load_dir(vimg, d.w, d.vn, x1, y1, v);
    const float2 hp = exact_hypothesis(c0, c1, d0, d1);
    const size_t o = ((size_t)job * d.vn + v) * d.hn + h;
    ws.hyp_true[o] = hp;
    const int kind = classify_hypothesis(hp.x, hp.y, fc.fast_ok != 0);
    const float qnan = __int_as_float(0x7fc00000);
    ws.hyp_filt[o] = kind == 0 ? hp : make_float2(qnan, qnan);
    if (kind == 2) {
      const int slot = atomicAdd(&ws.n_exact[job * d.vn + v], 1);
      ws.exact_list[((size_t)job * d.vn + v) * d.hn + slot] = h;
      if (ws.stats) atomicAdd(&ws.stats[2], 1ull);
    }
    ws.counts[o] = 0;
    if (dbg_hyps) {
      float* dst = dbg_hyps + ((((size_t)job * d.max_iter + rnd) * d.hn + h) * d.vn + v) * 2;
      dst[0] = hp.x;
      dst[1] = hp.y;
    }
  }
}

// ------------------------------------------------------------------------------------ plan
// single block; tile_px = pixels per scoring work item
__global__ void __launch_bounds__(1024) k_plan(WS ws, Dims d, int tile_px) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ int swarp[32];
  __shared__ int srun;
  if (tid == 0) srun = 0;
  __syncthreads();
  for (int s = 0; s < d.J; s += 1024) {
    const int job = s + tid;
    int nt = 0;
    if (job < d.J && (ws.job_flags[job] & JOB_ACTIVE)) nt = (ws.job_tn[job] + tile_px - 1) / tile_px;
    const int cnt = nt * d.vn;
    int x = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) swarp[warp] = x;
    __syncthreads();
    int woff = 0;
    for (int k = 0; k < warp; ++k) woff += swarp[k];
    const int run = srun;
    int pos = run + woff + x - cnt;
    for (int t = 0; t < nt; ++t)
      for (int v = 0; v < d.vn; ++v) {
        if (pos < d.max_items) ws.items[pos] = make_int2(job, (v << 20) | t);
        ++pos;
      }
    __syncthreads();
    if (tid == 1023) srun = run + woff + x;
    __syncthreads();
  }
  if (tid == 0) {
    ws.ctrl[CTRL_NITEMS] = min(srun, d.max_items);
    ws.ctrl[CTRL_WORK] = 0;
    ws.ctrl[CTRL_NACTIVE] = 0;
  }
}

// ------------------------------------------------------------------------------------ K3
struct ScoreArgs {
  WS ws;
  Dims d;
  FilterConsts fc;
  const float* vertex;
};

constexpr int kPxPerWarp = 128;                          // pixels of one warp's subset
constexpr int kScoreTile = kScoreWarps * kPxPerWarp;     // pixels per work item (1024)

// Exact inlier count of one hypothesis over pixels [t0, t0+npx) of a job, whole warp cooperating.
__device__ __noinline__ int exact_count(const uint32_t* __restrict__ pix, const float* __restrict__ vimg, int w, int vn,
                                        int v, int t0, int npx, float hx, float hy, float thr) {
  const int lane = threadIdx.x & 31;
  int c = 0;
  for (int q = lane; q < npx; q += 32) {
    const uint32_t pk = pix[t0 + q];
    const int x = pk & 0xFFFFu, y = pk >> 16;
    const float2 dv = load_dir(vimg, w, vn, x, y, v);
    c += exact_inlier(hx, hy, (float)x + 0.5f, (float)y + 0.5f, dv.x, dv.y, exact_norm(dv.x, dv.y), thr) ? 1 : 0;
  }
  return __reduce_add_sync(0xffffffffu, c);
}

// Persistent grid; each block pulls (job, keypoint, 1024-pixel tile) items from a global counter.
//   * the block turns the tile's pixels into 6 filter coefficients each (shared memory, 24 B/pixel);
//   * warp w owns pixels [w*128, w*128+128) of the tile and sweeps the hypotheses in groups of 32*H:
//     every lane keeps H hypotheses and their 2*H private vote counters in registers, the pixel
//     coefficients arrive as broadcast LDS.128 + LDS.64 — no shuffles, ballots or atomics in the loop;
//   * inlier <=> sign bit of t_lo = |p| - a (NaN and +0 count as "not inlier"), the second counter
//     counts t_hi = t_lo - kappa |p|; a hypothesis whose two counters differ met an uncertain unit
//     and is re-counted with the exact predicate.
template <int H>
__global__ void __launch_bounds__(kScoreThreads) k_score(ScoreArgs a) {
  __shared__ float4 sA[kScoreTile];  // (cx, cy, D, -E)
  __shared__ float2 sB[kScoreTile];  // (-G, -H)
  __shared__ int s_item;
  __shared__ int s_weird;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int hn = a.d.hn;
  const int n_items = a.ws.ctrl[CTRL_NITEMS];
  const float nkappa = -a.fc.kappa;
  const int n_groups = (hn + 32 * H - 1) / (32 * H);
  for (;;) {
    __syncthreads();  // previous item fully consumed
    if (tid == 0) {
      s_item = atomicAdd(&a.ws.ctrl[CTRL_WORK], 1);
      s_weird = 0;
    }
    __syncthreads();
    const int item = s_item;
    if (item >= n_items) break;
    const int2 it = a.ws.items[item];
    const int job = it.x, v = it.y >> 20, tile = it.y & 0xFFFFF;
    const int img = job / a.d.oc;
    const int tn = a.ws.job_tn[job];
    const uint32_t* pix = a.ws.pix + (size_t)img * a.d.cap + a.ws.job_off[job];
    const float* vimg = a.vertex + (size_t)img * a.d.hw * a.d.vn * 2;
    const int tile0 = tile * kScoreTile;

    bool weird = false;
#pragma unroll
    for (int i = tid; i < kScoreTile; i += kScoreThreads) {
      PixCoef pc;
      pc.cx = pc.cy = pc.D = pc.E = pc.G = pc.H = 0.f;
      const int t = tile0 + i;
      if (t < tn) {
        const uint32_t pk = pix[t];
        const int x = pk & 0xFFFFu, y = pk >> 16;
        const float2 dv = load_dir(vimg, a.d.w, a.d.vn, x, y, v);
        weird |= !make_coef((float)x + 0.5f, (float)y + 0.5f, dv.x, dv.y, a.fc.k_lo, pc);
      }
      sA[i] = make_float4(pc.cx, pc.cy, pc.D, -pc.E);
      sB[i] = make_float2(-pc.G, -pc.H);
    }
    if (weird) s_weird = 1;
    __syncthreads();
    const bool exact_tile = (a.fc.fast_ok == 0) || (s_weird != 0);

    const int t0 = tile0 + warp * kPxPerWarp;        // this warp's pixel subset
    const int npx = min(kPxPerWarp, tn - t0);        // warp-uniform; <= 0 -> nothing to do
    if (npx <= 0) continue;
    const size_t hoff = ((size_t)job * a.d.vn + v) * hn;
    int* gc = a.ws.counts + hoff;
    const float2* htrue = a.ws.hyp_true + hoff;

    if (exact_tile) {
      for (int h = 0; h < hn; ++h) {
        const float2 hp = htrue[h];
        const int c = exact_count(pix, vimg, a.d.w, a.d.vn, v, t0, npx, hp.x, hp.y, a.fc.thr);
        if (lane == 0 && c) atomicAdd(&gc[h], c);
      }
      if (a.ws.stats && lane == 0) atomicAdd(&a.ws.stats[3], 1ull);
      continue;
    }

    const float2* hfilt = a.ws.hyp_filt + hoff;
    const float4* cA = sA + warp * kPxPerWarp;
    const float2* cB = sB + warp * kPxPerWarp;
    for (int g = 0; g < n_groups; ++g) {
      float hx[H], hy[H];
      unsigned nlo[H], nhi[H];
#pragma unroll
      for (int i = 0; i < H; ++i) {
        const int h = (g * H + i) * 32 + lane;
        float2 hp = make_float2(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000));
        if (h < hn) hp = hfilt[h];
        hx[i] = hp.x;
        hy[i] = hp.y;
        nlo[i] = 0u;
        nhi[i] = 0u;
      }
#pragma unroll 2
      for (int q = 0; q < npx; ++q) {
        const float4 A = cA[q];
        const float2 B = cB[q];
#pragma unroll
        for (int i = 0; i < H; ++i) {
          const float hdx = hx[i] - A.x;  // the reference's rounded difference (:236)
          const float hdy = hy[i] - A.y;
          const float pv = fmaf(A.z, hdy, __fmul_rn(A.w, hdx));            // d^ x hd
          const float tlo = fmaf(B.x, hdx, fmaf(B.y, hdy, fabsf(pv)));      // |p| - k_lo d^.hd
          const float thi = fmaf(nkappa, fabsf(pv), tlo);                   // |p|/rho - k_lo d^.hd
          nlo[i] += __float_as_uint(tlo) >> 31;
          nhi[i] += __float_as_uint(thi) >> 31;
        }
      }
      // hypotheses that met an uncertain unit: exact re-count over this warp's pixels
#pragma unroll
      for (int i = 0; i < H; ++i) {
        unsigned m = __ballot_sync(0xffffffffu, nlo[i] != nhi[i]);
        while (m) {
          const int src = __ffs(m) - 1;
          m &= m - 1;
          const float bx = __shfl_sync(0xffffffffu, hx[i], src);
          const float by = __shfl_sync(0xffffffffu, hy[i], src);
          const int c = exact_count(pix, vimg, a.d.w, a.d.vn, v, t0, npx, bx, by, a.fc.thr);
          if (lane == src) nlo[i] = (unsigned)c;
          if (a.ws.stats && lane == 0) atomicAdd(&a.ws.stats[1], (unsigned long long)npx);
        }
      }
#pragma unroll
      for (int i = 0; i < H; ++i) {
        const int h = (g * H + i) * 32 + lane;
        if (h < hn && nlo[i]) atomicAdd(&gc[h], (int)nlo[i]);
      }
    }
    // hypotheses the filter cannot take (exact list)
    const int nlist = a.ws.n_exact[job * a.d.vn + v];
    for (int k = 0; k < nlist; ++k) {
      const int h = a.ws.exact_list[hoff + k];
      const float2 hp = htrue[h];
      const int c = exact_count(pix, vimg, a.d.w, a.d.vn, v, t0, npx, hp.x, hp.y, a.fc.thr);
      if (lane == 0 && c) atomicAdd(&gc[h], c);
    }
  }
}

// ------------------------------------------------------------------------------------ stop test
__device__ __forceinline__ double ipow_f64(double x, int n) {  // oracle/ransac_voting_np.py:_ipow_f64
  double r = 1.0, bs = x;
  while (n) {
    if (n & 1) r = __dmul_rn(r, bs);
    bs = __dmul_rn(bs, bs);
    n >>= 1;
  }
  return r;
}

// 1 - (1 - r^2)^hyp_num > confidence   (:344-346, all float32 in the reference)
__device__ __forceinline__ bool stop_test(float min_ratio, int hyp_num, float confidence) {
  const float r2 = __fmul_rn(min_ratio, min_ratio);
  const float base = __fsub_rn(1.0f, r2);
  const float pw = __double2float_rn(ipow_f64((double)base, hyp_num));
  return __fsub_rn(1.0f, pw) > confidence;
}

// ------------------------------------------------------------------------------------ K3b
// one block per job
__global__ void __launch_bounds__(256) k_update(WS ws, Dims d, int rnd, casa_ransac_debug dbg) {
  const int job = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int flags = ws.job_flags[job];
  if (!(flags & JOB_ACTIVE)) return;
  __shared__ unsigned long long sbest[8];
  __shared__ unsigned long long svbest[16];
  const int tn = ws.job_tn[job];
  for (int v = 0; v < d.vn; ++v) {
    const int* c = ws.counts + ((size_t)job * d.vn + v) * d.hn;
    unsigned long long best = 0ull;  // (count << 32) | ~h : max count, then lowest h (:328 argmax takes the first)
    for (int h = tid; h < d.hn; h += 256) {
      const unsigned long long key = ((unsigned long long)(unsigned)c[h] << 32) | (unsigned)(0xFFFFFFFFu - (unsigned)h);
      best = key > best ? key : best;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const unsigned long long y = __shfl_xor_sync(0xffffffffu, best, o);
      best = y > best ? y : best;
    }
    if (lane == 0) sbest[warp] = best;
    __syncthreads();
    if (tid == 0) {
      unsigned long long bb = sbest[0];
      for (int k = 1; k < 8; ++k) bb = sbest[k] > bb ? sbest[k] : bb;
      svbest[v] = bb;
    }
    __syncthreads();
  }
  if (dbg.counts) {
    int32_t* dst = dbg.counts + ((size_t)job * d.max_iter + rnd) * d.hn * d.vn;
    for (int e = tid; e < d.hn * d.vn; e += 256) {
      const int h = e / d.vn, v = e - h * d.vn;
      dst[e] = ws.counts[((size_t)job * d.vn + v) * d.hn + h];
    }
  }
  if (tid == 0) {
    float min_ratio = 3.0e38f;
    for (int v = 0; v < d.vn; ++v) {
      const int cnt = (int)(svbest[v] >> 32);
      const int widx = (int)(0xFFFFFFFFu - (unsigned)(svbest[v] & 0xFFFFFFFFull));
      const float ratio = __fdiv_rn((float)cnt, (float)tn);  // :333
      float best_ratio = ws.win_ratio[job * d.vn + v];
      if (best_ratio < ratio) {                              // :336-338
        best_ratio = ratio;
        ws.win_ratio[job * d.vn + v] = ratio;
        ws.win_pts[job * d.vn + v] = ws.hyp_true[((size_t)job * d.vn + v) * d.hn + widx];
      }
      min_ratio = fminf(min_ratio, best_ratio);              // :342
      if (dbg.win_idx) dbg.win_idx[((size_t)job * d.max_iter + rnd) * d.vn + v] = widx;
    }
    const int cur_iter = rnd + 1;                            // :341
    const int hyp_num = d.hn * cur_iter;                     // :340 (exact in float32 below 2^24)
    const bool stop = stop_test(min_ratio, hyp_num, d.confidence) || cur_iter >= d.max_iter;  // :344-347
    ws.job_rounds[job] = cur_iter;
    if (stop) {
      ws.job_flags[job] = flags & ~JOB_ACTIVE;
    } else {
      atomicAdd(&ws.ctrl[CTRL_NACTIVE], 1);
    }
    if (ws.stats) atomicAdd(&ws.stats[0], (unsigned long long)tn * d.vn * d.hn);
  }
}

// ------------------------------------------------------------------------------------ K4
// grid (vn, J), one block per (job, keypoint): re-vote the winner and accumulate the normal equations.
__global__ void __launch_bounds__(256) k_refine(WS ws, Dims d, FilterConsts fc, const float* __restrict__ vertex) {
  const int v = blockIdx.x, job = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int flags = ws.job_flags[job];
  const int tn = ws.job_tn[job];
  if ((flags & JOB_GATED) || tn <= 0) return;
  const int img = job / d.oc;
  const uint32_t* pix = ws.pix + (size_t)img * d.cap + ws.job_off[job];
  const float* vimg = vertex + (size_t)img * d.hw * d.vn * 2;
  const float2 wp = ws.win_pts[job * d.vn + v];
  double s[5] = {0, 0, 0, 0, 0};
  for (int t = tid; t < tn; t += 256) {
    const uint32_t pk = pix[t];
    const int x = pk & 0xFFFFu, y = pk >> 16;
    const float2 dv = load_dir(vimg, d.w, d.vn, x, y, v);
    const float cx = (float)x + 0.5f, cy = (float)y + 0.5f;
    if (exact_inlier(wp.x, wp.y, cx, cy, dv.x, dv.y, exact_norm(dv.x, dv.y), fc.thr)) {  // :353
      const float nx = __fmul_rn(dv.y, -1.0f);  // normal = (-dy, dx)                    :349
      const float ny = dv.x;
      const float bb = __fadd_rn(__fmul_rn(nx, cx), __fmul_rn(ny, cy));  // :359
      s[0] += (double)__fmul_rn(nx, nx);  // :361 float32 products, float64 accumulation
      s[1] += (double)__fmul_rn(nx, ny);
      s[2] += (double)__fmul_rn(ny, ny);
      s[3] += (double)__fmul_rn(nx, bb);  // :362
      s[4] += (double)__fmul_rn(ny, bb);
    }
  }
  __shared__ double sred[8][5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
#pragma unroll
    for (int o = 16; o; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
    if (lane == 0) sred[warp][k] = s[k];
  }
  __syncthreads();
  if (tid < 5) {
    double t = 0;
    for (int k = 0; k < 8; ++k) t += sred[k][tid];
    ws.sums[((size_t)job * d.vn + v) * 5 + tid] = t;
  }
}

// condition number of [[a,b],[b,c]], closed form in float64 (oracle: cond_2x2_sym_f64)
__device__ __forceinline__ bool invertible_2x2(float af, float bf, float cf) {
  const double a = af, b = bf, c = cf;
  const double m = __dmul_rn(__dadd_rn(a, c), 0.5);
  const double dd = __dmul_rn(__dsub_rn(a, c), 0.5);
  const double r = __dsqrt_rn(__dadd_rn(__dmul_rn(dd, dd), __dmul_rn(b, b)));
  double s0 = fabs(__dadd_rn(m, r));
  double s1 = fabs(__dsub_rn(m, r));
  if (s1 > s0) {
    const double t = s0;
    s0 = s1;
    s1 = t;
  }
  if (s1 == 0.0) return false;  // inf or nan condition number
  const double cnd = __ddiv_rn(s0, s1);
  return isfinite(cnd) && cnd < 1000000.0;  // :270-272, eps_inv = float32(1/1e-6)
}

// one thread per job
__global__ void __launch_bounds__(128) k_solve(WS ws, Dims d, float* __restrict__ out, casa_ransac_debug dbg) {
  const int job = blockIdx.x * 128 + threadIdx.x;
  if (job >= d.J) return;
  const int flags = ws.job_flags[job];
  const int tn = ws.job_tn[job];
  float2* o = reinterpret_cast<float2*>(out) + (size_t)job * d.vn;
  const bool dead = (flags & JOB_GATED) || tn <= 0;
  bool all_inv = !dead;
  if (!dead) {
    for (int v = 0; v < d.vn; ++v) {
      const double* s = ws.sums + ((size_t)job * d.vn + v) * 5;
      const float a = (float)s[0], b = (float)s[1], c = (float)s[2];
      all_inv = all_inv && invertible_2x2(a, b, c);
      if (dbg.ata) {
        float* q = dbg.ata + ((size_t)job * d.vn + v) * 3;
        q[0] = a; q[1] = b; q[2] = c;
      }
      if (dbg.atb) {
        float* q = dbg.atb + ((size_t)job * d.vn + v) * 2;
        q[0] = (float)s[3]; q[1] = (float)s[4];
      }
    }
  }
  for (int v = 0; v < d.vn; ++v) {
    float2 r = make_float2(0.f, 0.f);  // :291-292
    if (!dead) {
      r = ws.win_pts[job * d.vn + v];  // :364-365
      if (all_inv) {                   // :367  inv(ATA) @ ATb, closed form in float64 on the float32 sums
        const double* s = ws.sums + ((size_t)job * d.vn + v) * 5;
        const double a = (float)s[0], b = (float)s[1], c = (float)s[2], g0 = (float)s[3], g1 = (float)s[4];
        const double det = __dsub_rn(__dmul_rn(a, c), __dmul_rn(b, b));
        r.x = (float)__ddiv_rn(__dsub_rn(__dmul_rn(c, g0), __dmul_rn(b, g1)), det);
        r.y = (float)__ddiv_rn(__dsub_rn(__dmul_rn(a, g1), __dmul_rn(b, g0)), det);
      }
    }
    o[v] = r;
    if (dbg.win_pts) reinterpret_cast<float2*>(dbg.win_pts)[(size_t)job * d.vn + v] = dead ? make_float2(0.f, 0.f) : ws.win_pts[job * d.vn + v];
    if (dbg.win_ratio) dbg.win_ratio[(size_t)job * d.vn + v] = dead ? 0.f : ws.win_ratio[job * d.vn + v];
  }
  if (dbg.refined) dbg.refined[job] = (!dead && all_inv) ? 1 : 0;
  if (dbg.tn0) dbg.tn0[job] = ws.job_tn0[job];
  if (dbg.tn) dbg.tn[job] = (flags & JOB_GATED) ? 0 : tn;
  if (dbg.rounds) dbg.rounds[job] = dead ? 0 : ws.job_rounds[job];
  if (dbg.pix_off) dbg.pix_off[job] = ws.job_off[job];
}

}  // namespace casa
