// prio_probe — does a short high-priority kernel overtake the pending blocks of a long low-priority grid on B200?
// Design exploration for the lanes of casa_ransac_vote (not part of the library, the tests or the bench).
//   A: 6000 blocks x 256 threads, ~60 us each (a non-persistent scoring grid: ~10 waves on 148 SMs x 4 blocks)
//   B: 600 blocks x 256 threads, ~3 us each, enqueued 100 us after A started
// Prints when B started and finished relative to A's start, for: plain streams of equal priority, B's stream at the
// highest priority, graphs whose B node carries cudaKernelNodeAttributePriority, graphs launched into a
// high-priority stream.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o prio_probe prio_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      return 1;                                                                        \
    }                                                                                  \
  } while (0)

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// regs: 64 per thread via launch bounds, so that 4 blocks fill an SM's register file like k_score
__global__ void __launch_bounds__(256, 4) k_long(unsigned long long* stamp, int ns, float* sink) {
  const unsigned long long t0 = gtime();
  if (threadIdx.x == 0) atomicMin(&stamp[0], t0);
  float acc[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) acc[i] = (float)(threadIdx.x + i);
  while (gtime() - t0 < (unsigned long long)ns) {
#pragma unroll
    for (int k = 0; k < 16; ++k)
#pragma unroll
      for (int i = 0; i < 24; ++i) acc[i] = fmaf(acc[i], 1.0001f, 0.5f);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) s += acc[i];
  if (s == 12345.678f) sink[0] = s;
  if (threadIdx.x == 0) atomicMax(&stamp[1], gtime());
}

__global__ void __launch_bounds__(256) k_short(unsigned long long* stamp, int ns) {
  const unsigned long long t0 = gtime();
  if (threadIdx.x == 0) atomicMin(&stamp[2], t0);
  while (gtime() - t0 < (unsigned long long)ns) {
  }
  if (threadIdx.x == 0) atomicMax(&stamp[3], gtime());
}

__global__ void k_delay(int ns) {
  const unsigned long long t0 = gtime();
  while (gtime() - t0 < (unsigned long long)ns) {
  }
}

static int report(const char* name, unsigned long long* dstamp) {
  unsigned long long st[4];
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(st, dstamp, sizeof(st), cudaMemcpyDeviceToHost));
  printf("%-64s A: 0 .. %7.1f us   B: %7.1f .. %7.1f us\n", name, (st[1] - st[0]) * 1e-3, ((double)st[2] - (double)st[0]) * 1e-3,
         ((double)st[3] - (double)st[0]) * 1e-3);
  return 0;
}

static int reset(unsigned long long* dstamp) {
  unsigned long long st[4] = {~0ull, 0ull, ~0ull, 0ull};
  CK(cudaMemcpy(dstamp, st, sizeof(st), cudaMemcpyHostToDevice));
  return 0;
}

int main() {
  unsigned long long* dstamp;
  float* sink;
  CK(cudaMalloc(&dstamp, 32));
  CK(cudaMalloc(&sink, 4));
  int least = 0, greatest = 0;
  CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
  printf("stream priority range: least %d greatest %d\n", least, greatest);
  cudaStream_t sa, sb, sb_hi;
  CK(cudaStreamCreateWithFlags(&sa, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithPriority(&sb_hi, cudaStreamNonBlocking, greatest));
  const int nA = 6000, nB = 600, nsA = 60000, nsB = 3000, delay = 100000;

  for (int rep = 0; rep < 2; ++rep) {
    // 1. plain streams, equal priority
    if (reset(dstamp)) return 1;
    k_long<<<nA, 256, 0, sa>>>(dstamp, nsA, sink);
    k_delay<<<1, 1, 0, sb>>>(delay);
    k_short<<<nB, 256, 0, sb>>>(dstamp, nsB);
    if (report("streams, equal priority", dstamp)) return 1;
    // 2. B's stream at the highest priority
    if (reset(dstamp)) return 1;
    k_long<<<nA, 256, 0, sa>>>(dstamp, nsA, sink);
    k_delay<<<1, 1, 0, sb_hi>>>(delay);
    k_short<<<nB, 256, 0, sb_hi>>>(dstamp, nsB);
    if (report("streams, B's stream at the highest priority", dstamp)) return 1;
    // 3. launch attribute priority on B in an equal-priority stream
    if (reset(dstamp)) return 1;
    k_long<<<nA, 256, 0, sa>>>(dstamp, nsA, sink);
    k_delay<<<1, 1, 0, sb>>>(delay);
    {
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(nB);
      cfg.blockDim = dim3(256);
      cfg.stream = sb;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributePriority;
      at[0].val.priority = greatest;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      CK(cudaLaunchKernelEx(&cfg, k_short, dstamp, nsB));
    }
    if (report("streams, launch attribute priority on B", dstamp)) return 1;
  }

  // graphs: gA = {A}, gB = {delay -> B}; B's node with / without the priority attribute; launched into plain / high streams
  for (int mode = 0; mode < 3; ++mode) {
    cudaGraph_t gA, gB;
    cudaGraphExec_t eA, eB;
    CK(cudaGraphCreate(&gA, 0));
    CK(cudaGraphCreate(&gB, 0));
    cudaGraphNode_t nA_, nD, nB_;
    cudaKernelNodeParams kp;
    int nsA_ = nsA, nsB_ = nsB, delay_ = delay;
    void* argsA[] = {&dstamp, &nsA_, &sink};
    void* argsB[] = {&dstamp, &nsB_};
    void* argsD[] = {&delay_};
    memset(&kp, 0, sizeof(kp));
    kp.func = (void*)k_long;
    kp.gridDim = dim3(nA);
    kp.blockDim = dim3(256);
    kp.kernelParams = argsA;
    CK(cudaGraphAddKernelNode(&nA_, gA, nullptr, 0, &kp));
    kp.func = (void*)k_delay;
    kp.gridDim = dim3(1);
    kp.blockDim = dim3(1);
    kp.kernelParams = argsD;
    CK(cudaGraphAddKernelNode(&nD, gB, nullptr, 0, &kp));
    kp.func = (void*)k_short;
    kp.gridDim = dim3(nB);
    kp.blockDim = dim3(256);
    kp.kernelParams = argsB;
    CK(cudaGraphAddKernelNode(&nB_, gB, &nD, 1, &kp));
    if (mode == 1) {
      cudaKernelNodeAttrValue av;
      memset(&av, 0, sizeof(av));
      av.priority = greatest;
      CK(cudaGraphKernelNodeSetAttribute(nB_, cudaKernelNodeAttributePriority, &av));
      CK(cudaGraphKernelNodeSetAttribute(nD, cudaKernelNodeAttributePriority, &av));
    }
    CK(cudaGraphInstantiate(&eA, gA, 0));
    CK(cudaGraphInstantiate(&eB, gB, 0));
    for (int rep = 0; rep < 2; ++rep) {
      if (reset(dstamp)) return 1;
      CK(cudaGraphLaunch(eA, sa));
      CK(cudaGraphLaunch(eB, mode == 2 ? sb_hi : sb));
      if (report(mode == 0 ? "graphs, no priorities" : mode == 1 ? "graphs, B's nodes with the priority attribute" : "graphs, B's graph launched into a high-priority stream", dstamp))
        return 1;
    }
  }
  return 0;
}
