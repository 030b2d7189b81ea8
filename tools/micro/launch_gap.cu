// Kernel-to-kernel dependency latency inside a CUDA graph on this GPU, with and without programmatic dependent launch:
// a chain of N kernels, each spinning for ~`spin` ns on every SM (148 CTAs x 256 threads), captured into one graph.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o launch_gap launch_gap.cu && ./launch_gap
#include <cstdio>
#include <cuda_runtime.h>

__global__ void spin_kernel(int* sink, long long spin_ns, int pdl) {
  if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  long long t = t0;
  while (t - t0 < spin_ns) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  if (pdl) asm volatile("griddepcontrol.launch_dependents;");
  if (threadIdx.x == 0 && blockIdx.x == 0) sink[0] += 1;
}

static float run(int n, long long spin, int pdl, int blocks, int* sink) {
  cudaStream_t st;
  cudaStreamCreate(&st);
  cudaGraph_t g;
  cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
  for (int i = 0; i < n; ++i) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks), cfg.blockDim = dim3(256), cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at, cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, spin_kernel, sink, spin, pdl);
  }
  cudaStreamEndCapture(st, &g);
  cudaGraphInstantiate(&ge, g, 0);
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  for (int w = 0; w < 3; ++w) cudaGraphLaunch(ge, st);
  cudaEventRecord(a, st);
  for (int r = 0; r < 10; ++r) cudaGraphLaunch(ge, st);
  cudaEventRecord(b, st);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
  cudaGraphExecDestroy(ge), cudaGraphDestroy(g), cudaStreamDestroy(st);
  return ms * 1e3f / (10 * n);
}

int main() {
  int* sink;
  cudaMalloc(&sink, 4);
  cudaMemset(sink, 0, 4);
  for (int blocks : {1, 148, 1184})
    for (long long spin : {0LL, 5000LL, 20000LL})
      for (int pdl : {0, 1})
        printf("blocks %5d spin %6lld ns pdl %d : %.2f us per kernel (overhead %.2f)\n", blocks, spin, pdl, run(20, spin, pdl, blocks, sink),
               run(20, spin, pdl, blocks, sink) - spin * 1e-3f);
  return 0;
}
