// Forward kernels outside the tensor-core path (forward_tc.cu): stand-alone hash/OneBlob encodings (the tcnn seam),
// depth sampling, compositing of caller-provided samples, and the loss statistics.
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// encodings (lower seam)
// ---------------------------------------------------------------------------------------------
// thread = (point, level), level fastest: a half-warp writes one point's 128-byte feature row.
__global__ void __launch_bounds__(256) encode_fwd_kernel(const __grid_constant__ DevPlan P, const float2* __restrict__ grid,
                                                         const float* __restrict__ x, int64_t n, float2* __restrict__ out) {
  __shared__ DevLevel s_lv[NRT_L];
  if (threadIdx.x < NRT_L) s_lv[threadIdx.x] = P.lv[threadIdx.x];
  __syncthreads();
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t pt = t >> 4;
  int l = (int)(t & 15);
  if (pt >= n) return;
  float x0 = __ldg(x + pt * 3), x1 = __ldg(x + pt * 3 + 1), x2 = __ldg(x + pt * 3 + 2);
  out[pt * NRT_L + l] = level_gather(s_lv[l], grid, x0, x1, x2);
}

__global__ void __launch_bounds__(256) oneblob_fwd_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  // thread = (point, dim)
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * 3) return;
  float bins[NRT_BINS];
  oneblob16(__ldg(x + t), bins);
  float4* o = reinterpret_cast<float4*>(out + t * NRT_BINS);
#pragma unroll
  for (int q = 0; q < 4; ++q) o[q] = make_float4(bins[4 * q], bins[4 * q + 1], bins[4 * q + 2], bins[4 * q + 3]);
}

__global__ void __launch_bounds__(256) oneblob_bwd_kernel(const float* __restrict__ x, int64_t n, const float* __restrict__ dout,
                                                          float* __restrict__ dx) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * 3) return;
  float xv = __ldg(x + t);
  // d bin_b / dx = -(pdf_w((b+1)/n - x) - pdf_w(b/n - x)), pdf_w = wrapped kernel
  auto wpdf = [](float tt) { return quartic_pdf(tt) + quartic_pdf(tt - 1.0f) + quartic_pdf(tt + 1.0f); };
  float left = wpdf(0.0f - xv);
  const float first = left;
  float acc = 0.f;
#pragma unroll
  for (int b = 0; b < NRT_BINS; ++b) {
    float right = (b == NRT_BINS - 1) ? first : wpdf((float)(b + 1) * (1.0f / NRT_BINS) - xv);
    acc = fmaf(-(right - left), __ldg(dout + t * NRT_BINS + b), acc);
    left = right;
  }
  dx[t] = acc;
}

// ---------------------------------------------------------------------------------------------
// depth sampling only (API parity with render_rays' first half; the render kernel fuses the same code)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sample_z_kernel(const __grid_constant__ DevPlan P, const float* __restrict__ target_d,
                                                       int64_t n_rays, const float* __restrict__ u, int perturb, uint64_t seed,
                                                       float* __restrict__ z_out) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  float* z = smem + warp * P.S;
  for (int64_t ray = (int64_t)blockIdx.x * wpb + warp; ray < n_rays; ray += (int64_t)gridDim.x * wpb) {
    warp_sample_z(P, __ldg(target_d + ray), u ? u + ray * P.S : nullptr, perturb, seed, ray, z, lane);
    for (int s = lane; s < P.S; s += 32) z_out[ray * P.S + s] = z[s];
    __syncwarp();
  }
}

// raw2outputs / sdf2weights on caller-provided raw and z (any sample count S <= NRT_SMAX)
__global__ void __launch_bounds__(256) composite_fwd_kernel(const __grid_constant__ DevPlan P, const float* __restrict__ raw_in,
                                                            const float* __restrict__ z_in, int64_t n_rays, int S,
                                                            const NrtRenderOut out) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  float* z = smem + warp * (S * 6);
  float* raw = z + S;
  for (int64_t ray = (int64_t)blockIdx.x * wpb + warp; ray < n_rays; ray += (int64_t)gridDim.x * wpb) {
    for (int s = lane; s < S; s += 32) z[s] = __ldg(z_in + ray * S + s);
    for (int i = lane; i < S * 5; i += 32) raw[i] = __ldg(raw_in + ray * S * 5 + i);
    __syncwarp();
    RayOut ro = warp_composite(P, S, raw, z, out.weights ? out.weights + ray * S : nullptr, lane);
    if (lane == 0) {
      if (out.rgb) {
        out.rgb[ray * 3 + 0] = ro.rgb[0];
        out.rgb[ray * 3 + 1] = ro.rgb[1];
        out.rgb[ray * 3 + 2] = ro.rgb[2];
      }
      if (out.depth) out.depth[ray] = ro.depth;
      if (out.depth_var) out.depth_var[ray] = ro.depth_var;
      if (out.acc) out.acc[ray] = ro.acc;
      if (out.disp) out.disp[ray] = ro.disp;
      if (out.uncert) out.uncert[ray] = ro.uncert;
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// loss statistics (JointEncodingNaruto.forward train branch + get_masks/get_sdf_loss).
// Deterministic: per-block partial sums in fp64, reduced by the last block to finish.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) loss_partial_kernel(const __grid_constant__ DevPlan P, const NrtRenderOut rend,
                                                           const float* __restrict__ target_rgb,
                                                           const float* __restrict__ target_d, int64_t n_rays,
                                                           double* __restrict__ block_part, unsigned int* __restrict__ counter,
                                                           double* __restrict__ stats) {
  const int S = P.S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  double acc[NRT_N_STATS];
#pragma unroll
  for (int i = 0; i < NRT_N_STATS; ++i) acc[i] = 0.0;
  float umin = 3.4e38f;
  const float tr = P.sc_trunc;   // truncation = trunc * sc_factor
  for (int64_t ray = (int64_t)blockIdx.x * wpb + warp; ray < n_rays; ray += (int64_t)gridDim.x * wpb) {
    const float td = __ldg(target_d + ray);
    const bool valid = td > 0.0f && td < P.depth_trunc;
    if (lane == 0) {
      acc[NRT_STAT_N_RAYS] += 1.0;
      float e0 = rend.rgb[ray * 3 + 0] - __ldg(target_rgb + ray * 3 + 0);
      float e1 = rend.rgb[ray * 3 + 1] - __ldg(target_rgb + ray * 3 + 1);
      float e2 = rend.rgb[ray * 3 + 2] - __ldg(target_rgb + ray * 3 + 2);
      acc[NRT_STAT_RGB_SQ] += (double)(e0 * e0) + (double)(e1 * e1) + (double)(e2 * e2);
      float U = rend.uncert[ray];
      umin = fminf(umin, U);
      if (valid) {
        float e = rend.depth[ray] - td;
        acc[NRT_STAT_N_VALID] += 1.0;
        acc[NRT_STAT_DEPTH_SQ] += (double)(e * e);
        acc[NRT_STAT_INV2U] += (double)(1.0f / (2.0f * (U + 1e-9f)));
        acc[NRT_STAT_LOGU] += (double)logf(U + 1e-9f);
      }
    }
    // per-sample masks: front = z < d - tr ; back = z > d + tr ; sdf_mask = !front & !back & (d > 0)
    const float lo = __fsub_rn(td, tr), hi = __fadd_rn(td, tr);
    for (int s = lane; s < S; s += 32) {
      float zz = rend.z_vals[ray * S + s];
      float sdf = rend.raw[(ray * S + s) * 5 + 3];
      acc[NRT_STAT_N_SAMPLES] += 1.0;
      if (zz < lo) {
        float e = sdf - 1.0f;
        acc[NRT_STAT_N_FS] += 1.0;
        acc[NRT_STAT_FS_SQ] += (double)(e * e);
      } else if (!(zz > hi) && td > 0.0f) {
        float e = __fadd_rn(zz, __fmul_rn(sdf, tr)) - td;
        acc[NRT_STAT_N_SDF] += 1.0;
        acc[NRT_STAT_SDF_SQ] += (double)(e * e);
      }
    }
  }
  __shared__ double s_part[8][NRT_N_STATS];
  __shared__ float s_min[8];
  __shared__ bool s_last;
#pragma unroll
  for (int i = 0; i < NRT_N_STATS_SUM; ++i) {
    double v = warp_sum_d(acc[i]);
    if (lane == 0) s_part[warp][i] = v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) umin = fminf(umin, __shfl_xor_sync(0xffffffffu, umin, o));
  if (lane == 0) s_min[warp] = umin;
  __syncthreads();
  if (threadIdx.x < NRT_N_STATS) {
    double v = 0.0;
    if (threadIdx.x < NRT_N_STATS_SUM) {
      for (int w = 0; w < wpb; ++w) v += s_part[w][threadIdx.x];
    } else if (threadIdx.x == NRT_STAT_UNCERT_MIN) {
      float m = 3.4e38f;
      for (int w = 0; w < wpb; ++w) m = fminf(m, s_min[w]);
      v = (double)m;
    }
    block_part[(int64_t)blockIdx.x * NRT_N_STATS + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (s_last && threadIdx.x < NRT_N_STATS) {
    double v = threadIdx.x == NRT_STAT_UNCERT_MIN ? 3.4e38 : 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) {
      double pv = block_part[(int64_t)b * NRT_N_STATS + threadIdx.x];
      v = threadIdx.x == NRT_STAT_UNCERT_MIN ? fmin(v, pv) : v + pv;
    }
    stats[threadIdx.x] = v;
    if (threadIdx.x == 0) *counter = 0u;   // re-arm for the next launch
  }
}

__global__ void loss_finalize_kernel(const double* __restrict__ stats, float* __restrict__ losses) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  finalize_losses(stats, losses);
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
static inline int grid_for(int64_t work_items, int per_block, int sm_count, int waves_cap) {
  int64_t b = (work_items + per_block - 1) / per_block;
  int64_t cap = (int64_t)sm_count * waves_cap;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

int launch_encode_fwd(const NrtPlan* plan, const float* grid, const float* x, int64_t n, float* out, cudaStream_t st) {
  if (n == 0) return NRT_OK;
  int64_t threads = n * NRT_L;
  encode_fwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(plan->dev, (const float2*)grid, x, n, (float2*)out);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}

int launch_oneblob_fwd(const float* x, int64_t n, float* out, cudaStream_t st) {
  if (n == 0) return NRT_OK;
  oneblob_fwd_kernel<<<(unsigned)((n * 3 + 255) / 256), 256, 0, st>>>(x, n, out);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}

int launch_oneblob_bwd(const float* x, int64_t n, const float* dout, float* dx, cudaStream_t st) {
  if (n == 0) return NRT_OK;
  oneblob_bwd_kernel<<<(unsigned)((n * 3 + 255) / 256), 256, 0, st>>>(x, n, dout, dx);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}

int launch_sample_z(const NrtPlan* plan, const float* target_d, int64_t n_rays, const float* u, int perturb, uint64_t seed,
                    float* z, cudaStream_t st) {
  if (n_rays == 0) return NRT_OK;
  const size_t smem = 8 * plan->dev.S * sizeof(float);
  sample_z_kernel<<<grid_for(n_rays, 8, plan->sm_count, 8), 256, smem, st>>>(plan->dev, target_d, n_rays, u, perturb, seed, z);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}

int launch_composite_fwd(const NrtPlan* plan, const float* raw, const float* z, int64_t n_rays, int S, const NrtRenderOut* out,
                         cudaStream_t st) {
  if (n_rays == 0) return NRT_OK;
  const size_t smem = (size_t)8 * S * 6 * sizeof(float);
  composite_fwd_kernel<<<grid_for(n_rays, 8, plan->sm_count, 8), 256, smem, st>>>(plan->dev, raw, z, n_rays, S, *out);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}

// scratch for the deterministic loss reduction lives behind stats: [stats(16) | counter(2 doubles) | block partials]
#define LOSS_MAX_BLOCKS 296
int64_t loss_stats_doubles() { return NRT_N_STATS + 2 + (int64_t)LOSS_MAX_BLOCKS * NRT_N_STATS; }

int launch_loss_partial(const NrtPlan* plan, const NrtRenderOut* rend, const float* target_rgb, const float* target_d,
                        int64_t n_rays, double* stats, cudaStream_t st) {
  NRT_REQUIRE(rend->rgb && rend->depth && rend->uncert && rend->z_vals && rend->raw, "loss needs rgb, depth, uncert, z_vals, raw");
  unsigned int* counter = reinterpret_cast<unsigned int*>(stats + NRT_N_STATS);
  double* part = stats + NRT_N_STATS + 2;
  int blocks = grid_for(n_rays, 8, plan->sm_count, 2);
  if (blocks > LOSS_MAX_BLOCKS) blocks = LOSS_MAX_BLOCKS;
  loss_partial_kernel<<<blocks, 256, 0, st>>>(plan->dev, *rend, target_rgb, target_d, n_rays, part, counter, stats);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}

int launch_loss_finalize(const double* stats, float* losses, cudaStream_t st) {
  loss_finalize_kernel<<<1, 32, 0, st>>>(stats, losses);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}
