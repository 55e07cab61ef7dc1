// Micro-benchmark: per-kernel cost of a dependent chain inside a CUDA graph, with and without
// programmatic dependent launch (griddepcontrol).  nvcc -arch=sm_100a -o pdl_test pdl_test.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_step(float* x, int n, int pdl) {
  if (pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  __shared__ float s[32];
  if (threadIdx.x < 32) s[threadIdx.x] = threadIdx.x;   // some independent prologue
  __syncthreads();
  if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = x[(i + 1) % n] * 0.999f + s[threadIdx.x & 31] * 1e-9f;   // depends on the previous kernel
}
static float run(int pdl, int chain, int grid, int reps) {
  float* x; cudaMalloc(&x, grid * 256 * 4); cudaMemset(x, 0, grid * 256 * 4);
  cudaStream_t s; cudaStreamCreate(&s);
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < chain; ++i) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = 256; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    int n = grid * 256;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_step, x, n, pdl);
    if (e != cudaSuccess) { printf("launch err %s\n", cudaGetErrorString(e)); return -1; }
  }
  cudaError_t e = cudaStreamEndCapture(s, &g);
  if (e != cudaSuccess) { printf("capture err %s\n", cudaGetErrorString(e)); return -1; }
  e = cudaGraphInstantiate(&ge, g, 0);
  if (e != cudaSuccess) { printf("instantiate err %s\n", cudaGetErrorString(e)); return -1; }
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 5; ++i) cudaGraphLaunch(ge, s);
  cudaStreamSynchronize(s);
  cudaEventRecord(a, s);
  for (int i = 0; i < reps; ++i) cudaGraphLaunch(ge, s);
  cudaEventRecord(b, s);
  cudaStreamSynchronize(s);
  float ms; cudaEventElapsedTime(&ms, a, b);
  e = cudaGetLastError();
  if (e != cudaSuccess) printf("err %s\n", cudaGetErrorString(e));
  return ms * 1e3f / (reps * chain);
}
int main() {
  for (int grid : {8, 148, 592, 2368})
    printf("grid %5d: plain %.2f us/kernel   pdl %.2f us/kernel\n", grid, run(0, 60, grid, 200), run(1, 60, grid, 200));
  return 0;
}
