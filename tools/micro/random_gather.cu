// What HBM3e delivers for RANDOM 8-byte reads (one 32-byte sector each) out of a table that does not fit L2 -- the access pattern
// of the hashed levels at T = 2^21 (154 MB table, BASELINE configs[3]).  Every thread issues `ILP` independent loads per round
// from hashed indices; reports entries/s, sector GB/s (32 B per read) and useful GB/s (8 B per read).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o random_gather random_gather.cu && ./random_gather
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

template <int ILP>
__global__ void __launch_bounds__(256) gather_kernel(const float2* __restrict__ table, uint32_t mask, int rounds, float* __restrict__ out) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f;
  for (int r = 0; r < rounds; ++r) {
    float2 v[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) v[k] = __ldg(table + (mix(tid * 977u + r * 131071u + k * 7919u) & mask));
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc += v[k].x + v[k].y;
  }
  if (acc == 123.456f) out[0] = acc;
}

template <int ILP>
static void run(const float2* table, uint32_t mask, int blocks, float* out, const char* what) {
  const int rounds = 64;
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  gather_kernel<ILP><<<blocks, 256>>>(table, mask, 4, out);
  cudaEventRecord(a);
  gather_kernel<ILP><<<blocks, 256>>>(table, mask, rounds, out);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double n = (double)blocks * 256 * rounds * ILP;
  printf("%-28s blocks %5d ILP %2d: %7.3f ms  %6.1f G reads/s  sectors %7.1f GB/s  useful %6.1f GB/s\n", what, blocks, ILP, ms, n / ms * 1e-6,
         n * 32 / ms * 1e-6, n * 8 / ms * 1e-6);
}

int main() {
  float* out;
  cudaMalloc(&out, 4);
  for (int log2n : {16, 21, 24, 26}) {                  // entries of 8 bytes: 0.5 MB (L2), 16 MB (L2), 128 MB, 512 MB (HBM)
    const size_t n = (size_t)1 << log2n;
    float2* table;
    cudaMalloc(&table, n * sizeof(float2));
    cudaMemset(table, 0, n * sizeof(float2));
    char what[64];
    snprintf(what, sizeof what, "table %6.1f MB", n * 8 / 1e6);
    for (int blocks : {148 * 4, 148 * 8})
      run<16>(table, (uint32_t)(n - 1), blocks, out, what);
    run<32>(table, (uint32_t)(n - 1), 148 * 8, out, what);
    cudaFree(table);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
  return 0;
}
