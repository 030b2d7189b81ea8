// Goal-space uncertainty aggregation on the device (reference: NarutoPlanner.uncertainty_aggregation_v2,
// src/planner/naruto_planner.py:596-735) -- the consumer of the dense uncertainty / SDF sweep (nrt_map_volumes), so the
// volumes never leave HBM between mapping and planning ("planner hand-off on device").
//
// The reference materialises [G, k, 3] view vectors, a [G*k_valid, 30, 3] sample tensor and its SDF gather (GBs at
// G = 12 600 goal candidates x k = 300 targets).  Here: one CTA per goal candidate, one thread per target voxel:
//   distance mask (min < |goal - target| < max, in voxels)  ->  safety mask of the candidate (volume border, SDF of the
//   candidate and its six neighbours >= safe_sdf)  ->  visibility: 30 samples goal - t (goal - target), t = linspace(0, 1, 30),
//   truncated to voxel indices like .long(), all with SDF > 0  ->  collections[g, j] = uncert[target j] or 0, summed per goal.
// fp32 arithmetic in the reference's op order (sub, mul, norm), so every mask is bit-identical; the per-goal sum is a
// fixed-order tree (torch's reduction order differs in the last bits).
#include "common.cuh"

struct GoalSpec {
  int nx, ny, nz;         // volume dims
  int k;                  // targets
  float min_d, max_d;     // sensing range in voxels
  float safe_sdf;
};

__device__ __forceinline__ float vol_at(const float* __restrict__ v, const GoalSpec& s, int x, int y, int z) {
  return __ldg(v + ((int64_t)x * s.ny + y) * s.nz + z);
}

__global__ void __launch_bounds__(128) goal_aggregate_kernel(const float* __restrict__ uncert, const float* __restrict__ sdf,
                                                             const GoalSpec s, const float* __restrict__ goal_pts, int64_t G,
                                                             const float* __restrict__ topk, float* __restrict__ collections,
                                                             float* __restrict__ aggre, int* __restrict__ n_valid) {
  __shared__ float s_red[4];
  __shared__ int s_cnt[4];
  for (int64_t g = blockIdx.x; g < G; g += gridDim.x) {
    const float gx = __ldg(goal_pts + g * 3), gy = __ldg(goal_pts + g * 3 + 1), gz = __ldg(goal_pts + g * 3 + 2);
    const int ix = (int)gx, iy = (int)gy, iz = (int)gz;
    // candidate safety (naruto_planner.py:657-669): inside the volume with a one-voxel rim, SDF of the 7-stencil >= safe_sdf
    bool unsafe = ix < 1 || ix + 1 >= s.nx || iy < 1 || iy + 1 >= s.ny || iz < 1 || iz + 1 >= s.nz;
    {
      const int xm = max(ix - 1, 0), xp = min(ix + 1, s.nx - 1), ym = max(iy - 1, 0), yp = min(iy + 1, s.ny - 1),
                zm = max(iz - 1, 0), zp = min(iz + 1, s.nz - 1);
      unsafe = unsafe || vol_at(sdf, s, ix, iy, iz) < s.safe_sdf || vol_at(sdf, s, xp, iy, iz) < s.safe_sdf ||
               vol_at(sdf, s, xm, iy, iz) < s.safe_sdf || vol_at(sdf, s, ix, yp, iz) < s.safe_sdf ||
               vol_at(sdf, s, ix, ym, iz) < s.safe_sdf || vol_at(sdf, s, ix, iy, zp) < s.safe_sdf ||
               vol_at(sdf, s, ix, iy, zm) < s.safe_sdf;
    }
    float acc = 0.f;
    int cnt = 0;
    for (int j = threadIdx.x; j < s.k; j += blockDim.x) {
      const float tx = __ldg(topk + j * 3), ty = __ldg(topk + j * 3 + 1), tz = __ldg(topk + j * 3 + 2);
      const float vx = __fsub_rn(gx, tx), vy = __fsub_rn(gy, ty), vz = __fsub_rn(gz, tz);
      const float dist = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz)));
      bool ok = !unsafe && dist < s.max_d && dist > s.min_d;
      if (ok) {
        // visibility (naruto_planner.py:674-682): torch.linspace(0, 1, 30) in fp32
        const float step = 1.0f / 29.0f;
        float mn = 3.4e38f;
#pragma unroll 6
        for (int i = 0; i < 30; ++i) {
          const float t = i < 15 ? __fmul_rn(step, (float)i) : __fsub_rn(1.0f, __fmul_rn(step, (float)(29 - i)));
          const int px = (int)__fsub_rn(gx, __fmul_rn(t, vx)), py = (int)__fsub_rn(gy, __fmul_rn(t, vy)),
                    pz = (int)__fsub_rn(gz, __fmul_rn(t, vz));
          mn = fminf(mn, vol_at(sdf, s, px, py, pz));
        }
        ok = mn > 0.0f;
      }
      const float u = ok ? vol_at(uncert, s, (int)tx, (int)ty, (int)tz) : 0.f;
      collections[g * s.k + j] = u;
      acc += u;
      cnt += ok ? 1 : 0;
    }
    acc = warp_sum(acc);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0) {
      s_red[threadIdx.x >> 5] = acc;
      s_cnt[threadIdx.x >> 5] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      aggre[g] = (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
      const int c = s_cnt[0] + s_cnt[1] + s_cnt[2] + s_cnt[3];
      if (c && n_valid) atomicAdd(n_valid, c);
    }
    __syncthreads();
  }
}

int launch_goal_aggregate(const float* uncert, const float* sdf, const int* dims, const float* goal_pts, int64_t G,
                          const float* topk, int k, float min_d, float max_d, float safe_sdf, float* collections, float* aggre,
                          int* n_valid, int sm_count, cudaStream_t st) {
  if (n_valid) NRT_CUDA_CHECK(cudaMemsetAsync(n_valid, 0, sizeof(int), st));
  if (G == 0) return NRT_OK;
  GoalSpec s{dims[0], dims[1], dims[2], k, min_d, max_d, safe_sdf};
  const int64_t cap = (int64_t)sm_count * 16;
  const unsigned blocks = (unsigned)(G < cap ? G : cap);
  goal_aggregate_kernel<<<blocks, 128, 0, st>>>(uncert, sdf, s, goal_pts, G, topk, collections, aggre, n_valid);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}
