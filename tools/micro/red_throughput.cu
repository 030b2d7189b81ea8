// Microbenchmark: throughput of no-return global reductions (red.global.add.f32 / .v2.f32 / .v4.f32) per SM on B200, random
// 8/16-byte-aligned addresses inside an L2-resident table (6.5 MB, the size of the office0 hash table).  Decides whether
// pairing x-neighbour corners into one 16-byte reduction is worth it for the backward scatter (csrc/backward_q.cu).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_throughput red_throughput.cu && ./red_throughput
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s; }

template <int V>
__global__ void red_kernel(float* table, uint32_t n_slots, int iters, int active_lanes) {
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  const bool on = (int)(threadIdx.x & 31) < active_lanes;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint32_t slot = lcg(s) % n_slots;
      float* p = table + (size_t)slot * V;
      if (on) {
        if (V == 1) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(1.0f) : "memory");
        if (V == 2) asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(1.0f), "f"(2.0f) : "memory");
        if (V == 4) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(1.0f), "f"(2.0f), "f"(3.0f), "f"(4.0f) : "memory");
      }
    }
  }
}

template <int V>
void run(float* table, size_t table_floats, int warps_per_cta, int active_lanes) {
  const int iters = 200;
  const uint32_t n_slots = (uint32_t)(table_floats / V);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  red_kernel<V><<<148, warps_per_cta * 32>>>(table, n_slots, 10, active_lanes);
  cudaEventRecord(a);
  red_kernel<V><<<148, warps_per_cta * 32>>>(table, n_slots, iters, active_lanes);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double instr_per_sm = (double)warps_per_cta * iters * 8;
  const double lanes_per_sm = instr_per_sm * active_lanes;
  printf("V=%d warps/SM=%2d lanes=%2d: %8.3f ms  %7.1f ns/warp-instr/SM  %6.2f ns/lane/SM  (%.2f G lane-reds/s chip, %.1f GB/s payload)\n", V,
         warps_per_cta, active_lanes, ms, ms * 1e6 / instr_per_sm, ms * 1e6 / lanes_per_sm, lanes_per_sm * 148 / ms / 1e6,
         lanes_per_sm * 148 * V * 4 / ms / 1e6);
}

int main() {
  const size_t table_floats = 1628176;   // office0 table
  float* table;
  cudaMalloc(&table, table_floats * sizeof(float));
  cudaMemset(table, 0, table_floats * sizeof(float));
  for (int warps : {8, 16, 32}) {
    for (int lanes : {32, 16}) {
      run<1>(table, table_floats, warps, lanes);
      run<2>(table, table_floats, warps, lanes);
      run<4>(table, table_floats, warps, lanes);
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
