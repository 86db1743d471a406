#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
__global__ void k_body(int* ctr, cudaGraphConditionalHandle h) {
  int v = atomicAdd(ctr, 1);
  cudaGraphSetConditional(h, v < 4 ? 1u : 0u);
}
__global__ void k_first(int* ctr, cudaGraphConditionalHandle h) { *ctr = 0; cudaGraphSetConditional(h, 1u); }
int main() {
  setvbuf(stdout, NULL, _IONBF, 0);
  int* d; cudaMalloc(&d, 4);
  cudaGraph_t g; cudaGraphCreate(&g, 0);
  cudaGraphConditionalHandle h;
  cudaGraphConditionalHandleCreate(&h, g, 0, cudaGraphCondAssignDefault);
  cudaGraphNode_t n0, nw, nb;
  cudaKernelNodeParams kp = {};
  void* a0[] = {&d, &h};
  kp.func = (void*)k_first; kp.gridDim = 1; kp.blockDim = 1; kp.kernelParams = a0;
  printf("add0 %d\n", cudaGraphAddKernelNode(&n0, g, nullptr, 0, &kp));
  cudaGraphNodeParams cp = {};
  cp.type = cudaGraphNodeTypeConditional;
  cp.conditional.handle = h; cp.conditional.type = cudaGraphCondTypeWhile; cp.conditional.size = 1;
  printf("addw %d\n", cudaGraphAddNode(&nw, g, &n0, 1, &cp));
  if (!cp.conditional.phGraph_out) { printf("no body graph\n"); return 1; }
  cudaGraph_t body = cp.conditional.phGraph_out[0];
  printf("body %p\n", (void*)body);
  kp.func = (void*)k_body;
  printf("addb %d\n", cudaGraphAddKernelNode(&nb, body, nullptr, 0, &kp));
  cudaEvent_t ev; cudaEventCreate(&ev);
  cudaGraphNode_t ne;
  if (getenv("EV")) printf("event-in-body %d\n", cudaGraphAddEventRecordNode(&ne, body, &nb, 1, ev));
  if (getenv("MC")) { static int hbuf; cudaGraphNode_t nm; printf("memcpy-in-body %d\n", cudaGraphAddMemcpyNode1D(&nm, body, &nb, 1, d + 0, d + 0, 0 + 4, cudaMemcpyDeviceToDevice)); (void)hbuf; }
  cudaGraphExec_t e;
  { cudaError_t r = cudaGraphInstantiate(&e, g, 0); printf("inst %d\n", r); if (r) return 1; }
  // per-call update of a node inside the body
  printf("setparams body %d\n", cudaGraphExecKernelNodeSetParams(e, nb, &kp));
  printf("launch %d\n", cudaGraphLaunch(e, 0));
  printf("sync %d\n", cudaDeviceSynchronize());
  int hv; cudaMemcpy(&hv, d, 4, cudaMemcpyDeviceToHost); printf("ctr %d (expect 6)\n", hv);
  return 0;
}
