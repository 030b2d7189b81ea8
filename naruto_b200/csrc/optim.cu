// Smoothness regulariser (CoSLAM.smoothness, tp/coslam.py:245-269) and the fused Adam step
// (torch.optim.Adam as configured at src/slam/coslam/coslam.py:409-419, 240-243).
#include "common.cuh"

// lattice point (i,j,k) of the (n-1)^3 smoothness lattice, normalised to the bound.  Mirrors the reference's
// op order: pts = (coords + jitter) * voxel + bb_min + offset ; x = (pts - bb_min) / (bb_max - bb_min),
// offset = rand3 * (extent - (n-1)*voxel - 2*margin) + margin
struct LatticeSpec {
  int n;               // smooth_pts; the lattice has (n-1)^3 points
  float voxel;         // float32(smooth_vox)
  float grid_size;     // float32((n-1) * smooth_vox), product taken in double like the reference's python scalar
  float margin;        // float32(smooth_margin)
  float margin2;       // float32(2 * smooth_margin)
};

__device__ __forceinline__ float lattice_coord(const DevPlan& P, int d, int i, const float* __restrict__ rnd,
                                               const LatticeSpec& ls) {
  const float off_max = __fsub_rn(__fsub_rn(P.bb_ext[d], ls.grid_size), ls.margin2);
  const float offset = __fadd_rn(__fmul_rn(rnd[d], off_max), ls.margin);
  float p = __fmul_rn(__fadd_rn((float)i, rnd[3 + d]), ls.voxel);
  p = __fadd_rn(__fadd_rn(p, P.bb_min[d]), offset);
  return normalise1(P, d, p);
}

// Both passes run with blockIdx.y = level and thread = lattice point (z index fastest), so a warp holds 32 neighbouring
// points of ONE level: level constants are uniform (constant bank, no shared-memory table), neighbouring points share or
// adjoin cells (L1 hits on the gather, reductions into neighbouring sectors), and the feature workspace F[level][point] is
// read and written coalesced.  (The first version used thread = (point, level) with the level fastest: 16 different level
// tables per warp, a bank-conflicted shared-memory level table and a strided workspace -- 14 + 31 us per iteration.)

// pass 1: hash features of every lattice point -> F[16][m^3] (float2)
__global__ void __launch_bounds__(256) smooth_encode_kernel(const __grid_constant__ DevPlan P, const float2* __restrict__ grid,
                                                            const float* __restrict__ rnd, const LatticeSpec ls,
                                                            float2* __restrict__ F) {
  const int l = blockIdx.y;
  const int m = ls.n - 1;
  const int64_t total = (int64_t)m * m * m;
  const int64_t pt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pt >= total) return;
  const int k = (int)(pt % m), j = (int)((pt / m) % m), i = (int)(pt / ((int64_t)m * m));
  const float x0 = lattice_coord(P, 0, i, rnd, ls);
  const float x1 = lattice_coord(P, 1, j, rnd, ls);
  const float x2 = lattice_coord(P, 2, k, rnd, ls);
  F[(int64_t)l * total + pt] = level_gather(P.lv[l], grid, x0, x1, x2);
}

// pass 2: TV loss + its gradient scattered straight into the table.
// d/dF[p] of sum over edges (F[p]-F[q])^2 = 2 * sum over neighbours (F[p]-F[q]).
__global__ void __launch_bounds__(256) smooth_tv_bwd_kernel(const __grid_constant__ DevPlan P, const float* __restrict__ rnd,
                                                            const LatticeSpec ls, const float2* __restrict__ F,
                                                            float loss_scale, float* __restrict__ loss, float2* __restrict__ dgrid,
                                                            int64_t pt_lo, int64_t pt_hi) {
  __shared__ float s_red[8];
  const int l = blockIdx.y;
  const int n = ls.n;
  const int m = n - 1;
  const int64_t total = (int64_t)m * m * m;
  // this launch owns the lattice points [pt_lo, pt_hi) of the scan order (all of them on one GPU; a slab per rank otherwise:
  // every edge is counted at its lower end and every gradient belongs to one point, so the slabs add up to the whole term)
  const int64_t pt = pt_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const float2* __restrict__ Fl = F + (int64_t)l * total;
  float tv = 0.f;
  if (pt < pt_hi) {
    const int idx[3] = {(int)(pt / ((int64_t)m * m)), (int)((pt / m) % m), (int)(pt % m)};
    const int64_t stride[3] = {(int64_t)m * m, m, 1};
    const float2 f = Fl[pt];
    float2 g = make_float2(0.f, 0.f);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (idx[d] + 1 < m) {
        const float2 q = Fl[pt + stride[d]];
        const float ex = f.x - q.x, ey = f.y - q.y;
        tv += ex * ex + ey * ey;      // each edge counted once, at its lower end
        g.x += ex;
        g.y += ey;
      }
      if (idx[d] > 0) {
        const float2 q = Fl[pt - stride[d]];
        g.x += f.x - q.x;
        g.y += f.y - q.y;
      }
    }
    const float inv = 1.0f / ((float)n * (float)n * (float)n);
    g.x *= 2.0f * inv * loss_scale;
    g.y *= 2.0f * inv * loss_scale;
    if (dgrid && (g.x != 0.f || g.y != 0.f)) {
      const float x0 = lattice_coord(P, 0, idx[0], rnd, ls);
      const float x1 = lattice_coord(P, 1, idx[1], rnd, ls);
      const float x2 = lattice_coord(P, 2, idx[2], rnd, ls);
      const DevLevel& L = P.lv[l];
      uint32_t e[8];
      float w[8];
      level_corners(L, x0, x1, x2, e, w);
      float2* base = dgrid + L.offset;
#pragma unroll
      for (int c = 0; c < 8; c += 2)          // x-neighbour corners share one 16-byte reduction when aligned (common.cuh)
        red_add_xpair(base, e[c], e[c + 1], w[c] * g.x, w[c] * g.y, w[c + 1] * g.x, w[c + 1] * g.y);
    }
    tv *= inv;
  }
  tv = warp_sum(tv);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = tv;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += s_red[w];
    atomicAdd(loss, v);
  }
}

int launch_smooth(const NrtPlan* plan, const float* grid, const float* rnd6, int n, double voxel, double margin,
                  float loss_scale, float* loss, float* dgrid, void* workspace, int part, int n_parts, cudaStream_t st) {
  const int m = n - 1;
  const int64_t pts = (int64_t)m * m * m;
  const dim3 blocks((unsigned)((pts + 255) / 256), NRT_L);
  const int64_t lo = pts * part / n_parts, hi = pts * (part + 1) / n_parts;      // this part's slab of the scan order
  float2* F = reinterpret_cast<float2*>(workspace);
  LatticeSpec ls{n, (float)voxel, (float)((double)(n - 1) * voxel), (float)margin, (float)(2.0 * margin)};
  // The accumulator is cleared by a memset node, not by the first kernel: on the forked branch of the mapping iteration
  // (mapper.py) that node also lets the ray path's next kernel take its SMs first -- measured 205 -> 187 us per iteration at
  // 2048 rays x 43 samples against clearing it inside smooth_encode_kernel (profiles/r02d_smooth_branch.log).
  NRT_CUDA_CHECK(cudaMemsetAsync(loss, 0, sizeof(float), st));
  smooth_encode_kernel<<<blocks, 256, 0, st>>>(plan->dev, (const float2*)grid, rnd6, ls, F);
  NRT_CUDA_CHECK(cudaGetLastError());
  if (hi > lo) {
    const dim3 tv_blocks((unsigned)((hi - lo + 255) / 256), NRT_L);
    smooth_tv_bwd_kernel<<<tv_blocks, 256, 0, st>>>(plan->dev, rnd6, ls, F, loss_scale, loss, (float2*)dgrid, lo, hi);
  }
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}

// ---------------------------------------------------------------------------------------------
// Adam.  step_dev (optional) holds the 1-based step count on the device so that a captured CUDA graph can
// be replayed; otherwise `step` is used.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, int step, const int* __restrict__ step_dev,
                                                   float lr, float beta1, float beta2, float eps, float wd, int zero_grad) {
  __shared__ float s_c[2];
  if (threadIdx.x == 0) {
    const int st = step_dev ? *step_dev : step;
    const double bc1 = 1.0 - pow((double)beta1, (double)st);
    const double bc2 = 1.0 - pow((double)beta2, (double)st);
    s_c[0] = (float)((double)lr / bc1);          // step_size
    s_c[1] = (float)sqrt(bc2);                   // bias_correction2_sqrt
  }
  __syncthreads();
  const float step_size = s_c[0], bc2s = s_c[1];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float pi = p[i], gi = g[i], mi = m[i], vi = v[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    mi = mi + (gi - mi) * (1.0f - beta1);                     // exp_avg.lerp_(grad, 1-beta1)
    vi = vi * beta2 + (1.0f - beta2) * gi * gi;               // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1-beta2)
    const float denom = sqrtf(vi) / bc2s + eps;
    pi = pi - step_size * (mi / denom);
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
    if (zero_grad) g[i] = 0.f;
  }
}

// All parameter groups of the scene model in ONE launch (create_optimizer's two groups + the uncertainty grid's own Adam):
// the groups are ranges of the flat [grid | decoder | uncert] vectors; a disabled group (the uncertainty grid on four
// iterations out of five) is skipped and keeps accumulating its gradient.  Same arithmetic as adam_kernel.
struct AdamRange {
  int64_t begin, end;            // floats; begin is a multiple of 4
  float lr, beta1, beta2, eps, wd;
  const int* step_dev;
  int enabled;
};

__device__ __forceinline__ void adam_bias_terms(const AdamRange& R, float* out) {
  const int st = R.enabled ? *R.step_dev : 1;
  const double bc1 = 1.0 - pow((double)R.beta1, (double)st), bc2 = 1.0 - pow((double)R.beta2, (double)st);
  out[0] = (float)((double)R.lr / bc1);          // step_size
  out[1] = (float)sqrt(bc2);                     // bias_correction2_sqrt
}

__device__ __forceinline__ void adam_range(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                           const AdamRange& R, float step_size, float bc2s, int zero_grad) {
  const float beta2 = R.beta2, eps = R.eps, wd = R.wd, omb1 = 1.0f - R.beta1, omb2 = 1.0f - R.beta2;
  auto one = [&](float& pi, float gi, float& mi, float& vi) {
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    mi = mi + (gi - mi) * omb1;
    vi = vi * beta2 + omb2 * gi * gi;
    const float denom = sqrtf(vi) / bc2s + eps;
    pi = pi - step_size * (mi / denom);
  };
  const int t = threadIdx.x;
  const int64_t b4 = R.begin >> 2, e4 = R.end >> 2;                 // whole float4s, then a scalar tail
  for (int64_t i = b4 + (int64_t)blockIdx.x * blockDim.x + t; i < e4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pi = reinterpret_cast<float4*>(p)[i], mi = reinterpret_cast<float4*>(m)[i], vi = reinterpret_cast<float4*>(v)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    one(pi.x, gv.x, mi.x, vi.x);
    one(pi.y, gv.y, mi.y, vi.y);
    one(pi.z, gv.z, mi.z, vi.z);
    one(pi.w, gv.w, mi.w, vi.w);
    reinterpret_cast<float4*>(p)[i] = pi;
    reinterpret_cast<float4*>(m)[i] = mi;
    reinterpret_cast<float4*>(v)[i] = vi;
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (blockIdx.x == 0) {
    const int64_t i = (e4 << 2) + t;
    if (i < R.end) {
      float pi = p[i], mi = m[i], vi = v[i];
      one(pi, g[i], mi, vi);
      p[i] = pi, m[i] = mi, v[i] = vi;
      if (zero_grad) g[i] = 0.f;
    }
  }
}

__global__ void __launch_bounds__(256) adam_groups_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                          float* __restrict__ v, const __grid_constant__ AdamRange r0,
                                                          const __grid_constant__ AdamRange r1, const __grid_constant__ AdamRange r2,
                                                          int zero_grad) {
  __shared__ float s_c[3][2];
  const int t = threadIdx.x;
  if (t == 0) adam_bias_terms(r0, s_c[0]);
  if (t == 32) adam_bias_terms(r1, s_c[1]);
  if (t == 64) adam_bias_terms(r2, s_c[2]);
  __syncthreads();
  if (r0.enabled) adam_range(p, g, m, v, r0, s_c[0][0], s_c[0][1], zero_grad);
  if (r1.enabled) adam_range(p, g, m, v, r1, s_c[1][0], s_c[1][1], zero_grad);
  if (r2.enabled) adam_range(p, g, m, v, r2, s_c[2][0], s_c[2][1], zero_grad);
}

int launch_adam_groups(float* p, float* g, float* m, float* v, const NrtAdamGroup* groups, int n_groups, int zero_grad, int sm_count,
                       cudaStream_t st) {
  AdamRange r[3] = {};
  for (int i = 0; i < n_groups && i < 3; ++i) {
    r[i].begin = groups[i].begin, r[i].end = groups[i].end;
    r[i].lr = groups[i].lr, r[i].beta1 = groups[i].beta1, r[i].beta2 = groups[i].beta2, r[i].eps = groups[i].eps;
    r[i].wd = groups[i].weight_decay, r[i].step_dev = groups[i].step_dev, r[i].enabled = groups[i].enabled;
  }
  adam_groups_kernel<<<sm_count * 8, 256, 0, st>>>(p, g, m, v, r[0], r[1], r[2], zero_grad);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}

__global__ void counter_add_kernel(int* c, int delta) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *c += delta;
}

int launch_adam(float* p, float* g, float* m, float* v, int64_t n, int step, const int* step_dev, float lr, float b1, float b2,
                float eps, float wd, int zero_grad, int sm_count, cudaStream_t st) {
  if (n == 0) return NRT_OK;
  int64_t blocks = (n + 255) / 256;
  int64_t cap = (int64_t)sm_count * 8;
  if (blocks > cap) blocks = cap;
  adam_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, g, m, v, n, step, step_dev, lr, b1, b2, eps, wd, zero_grad);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}

// Start of a mapping iteration inside a replayed CUDA graph: advance the device-side step counter and draw the six uniforms
// of the smoothness lattice (torch.rand(3), torch.rand((1,1,1,3)) in tp/coslam.py:252-258) from Philox keyed by (seed, step),
// so no host value and no separate RNG launch is needed per iteration.
__global__ void step_begin_kernel(int* c, int delta, uint64_t seed, float* __restrict__ rand6, int* c2) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int step = *c + delta;
  *c = step;
  if (c2) *c2 += 1;            // second counter (the uncertainty grid's Adam step on the iterations that step it)
  if (rand6) {
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const uint4 a = philox4x32(make_uint4((uint32_t)step, 0x534d4f4fu, 0u, 0u), key);
    const uint4 b = philox4x32(make_uint4((uint32_t)step, 0x534d4f4fu, 1u, 0u), key);
    rand6[0] = u32_to_unit(a.x);
    rand6[1] = u32_to_unit(a.y);
    rand6[2] = u32_to_unit(a.z);
    rand6[3] = u32_to_unit(a.w);
    rand6[4] = u32_to_unit(b.x);
    rand6[5] = u32_to_unit(b.y);
  }
}

int launch_step_begin(int* c, int delta, uint64_t seed, float* rand6, int* c2, cudaStream_t st) {
  step_begin_kernel<<<1, 32, 0, st>>>(c, delta, seed, rand6, c2);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}

int launch_counter_add(int* c, int delta, cudaStream_t st) {
  counter_add_kernel<<<1, 32, 0, st>>>(c, delta);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}
