// Backward kernels of the mapping iteration:
//   composite_bwd_kernel : dL/d raw[B,S,5] from the loss definition (per ray, one warp)
//   (backward_tc.cu)     : tcgen05 MLP backward: dL/d raw -> dL/d hash-features, dW1..dW4, uncertainty-grid gradient
//   encode_bwd_kernel    : scatter dL/d hash-features into the table with vectorised reductions
// Replaces what autograd does for loss.backward() (src/slam/coslam/coslam.py:216,368) through
// src/slam/coslam/model/scene_rep.py and the tcnn grid backward.
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// dL/d raw from the losses of JointEncodingNaruto.forward
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) composite_bwd_kernel(const __grid_constant__ DevPlan P, const NrtRenderOut rend,
                                                            const float* __restrict__ target_rgb,
                                                            const float* __restrict__ target_d, int64_t n_rays,
                                                            const double* __restrict__ stats,
                                                            const float* __restrict__ loss_grad, float* __restrict__ draw) {
  extern __shared__ __align__(16) float smem[];
  const int S = P.S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  float* z = smem + warp * (S * 7);
  float* raw = z + S;
  float* wbuf = raw + S * 5;

  // global scalars (identical on every rank once stats are all-reduced)
  const float Bn = (float)stats[NRT_STAT_N_RAYS], V = (float)stats[NRT_STAT_N_VALID], NS = (float)stats[NRT_STAT_N_SAMPLES];
  const float nfs = (float)stats[NRT_STAT_N_FS], nsdf = (float)stats[NRT_STAT_N_SDF];
  const float fs_w = 1.0f - nfs / (nfs + nsdf), sdf_w = 1.0f - nsdf / (nfs + nsdf);
  const float depth_loss = (float)(stats[NRT_STAT_DEPTH_SQ] / stats[NRT_STAT_N_VALID]);
  const float mean_inv2u = (float)(stats[NRT_STAT_INV2U] / stats[NRT_STAT_N_VALID]);
  const float g_rgbl = loss_grad[NRT_LOSS_RGB], g_depl = loss_grad[NRT_LOSS_DEPTH], g_sdfl = loss_grad[NRT_LOSS_SDF];
  const float g_fsl = loss_grad[NRT_LOSS_FS], g_uncl = loss_grad[NRT_LOSS_UNCERT];
  const float tr = P.sc_trunc;
  const bool have_out = rend.rgb && rend.depth && rend.uncert;

  for (int64_t ray = (int64_t)blockIdx.x * wpb + warp; ray < n_rays; ray += (int64_t)gridDim.x * wpb) {
    for (int s = lane; s < S; s += 32) z[s] = rend.z_vals[ray * S + s];
    for (int i = lane; i < S * 5; i += 32) raw[i] = rend.raw[ray * S * 5 + i];
    __syncwarp();
    RayOut ro;
    if (have_out) {
      // the forward left the ray's composited colour / depth / uncertainty in its output buffers: only the weights are re-formed
      warp_weights(P, S, raw, z, wbuf, lane, ro.z_cut, ro.wsum);
      ro.rgb[0] = __ldg(rend.rgb + ray * 3), ro.rgb[1] = __ldg(rend.rgb + ray * 3 + 1), ro.rgb[2] = __ldg(rend.rgb + ray * 3 + 2);
      ro.depth = __ldg(rend.depth + ray);
      ro.uncert = __ldg(rend.uncert + ray);
    } else {
      ro = warp_composite(P, S, raw, z, wbuf, lane);
    }
    __syncwarp();
    const float td = __ldg(target_d + ray);
    const bool valid = td > 0.0f && td < P.depth_trunc;
    float g_rgb[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) g_rgb[c] = g_rgbl * 2.0f * (ro.rgb[c] - __ldg(target_rgb + ray * 3 + c)) / (3.0f * Bn);
    float g_depth = 0.f, g_U = 0.f;
    if (valid) {
      const float e = ro.depth - td;
      const float Ue = ro.uncert + 1e-9f;
      g_depth = (g_depl + g_uncl * mean_inv2u) * 2.0f * e / V;
      g_U = g_uncl * (-depth_loss / (2.0f * Ue * Ue) + 0.5f / Ue) / V;
    }
    // D = sum_s dL/dw_s * w_s
    float D = 0.f;
    // (samples outside the truncation window have w == 0: they add nothing to D and get no colour / uncertainty gradient,
    // so their activations are never evaluated -- ~85 % of the samples)
    for (int s = lane; s < S; s += 32) {
      const float w = wbuf[s];
      if (w == 0.f) continue;
      const float unc = softplusf_(raw[s * 5 + 4]) + 0.01f;
      float dw = g_rgb[0] * sigmoidf_(raw[s * 5]) + g_rgb[1] * sigmoidf_(raw[s * 5 + 1]) + g_rgb[2] * sigmoidf_(raw[s * 5 + 2]) +
                 g_depth * z[s] + g_U * 2.0f * w * unc;
      D = fmaf(dw, w, D);
    }
    D = warp_sum(D);
    const float inv_den = 1.0f / (ro.wsum + 1e-8f);
    const float lo = __fsub_rn(td, tr), hi = __fadd_rn(td, tr);
    for (int s = lane; s < S; s += 32) {
      const float w = wbuf[s];
      const float sdf = raw[s * 5 + 3], ur = raw[s * 5 + 4], zz = z[s];
      float g[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
      float dw = 0.f;
      if (w != 0.f) {
        const float unc = softplusf_(ur) + 0.01f;
        float c[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          c[k] = sigmoidf_(raw[s * 5 + k]);
          g[k] = g_rgb[k] * w * c[k] * (1.0f - c[k]);
        }
        dw = g_rgb[0] * c[0] + g_rgb[1] * c[1] + g_rgb[2] * c[2] + g_depth * zz + g_U * 2.0f * w * unc;
        g[4] = g_U * w * w * (ur > 20.0f ? 1.0f : sigmoidf_(ur));
      } else {
        // w == 0 but inside the window (bell weight underflowed): dL/dw still flows through the normalisation
        dw = g_depth * zz;
        if (zz < ro.z_cut) {
          dw += g_rgb[0] * sigmoidf_(raw[s * 5]) + g_rgb[1] * sigmoidf_(raw[s * 5 + 1]) + g_rgb[2] * sigmoidf_(raw[s * 5 + 2]);
        }
      }
      float gs = 0.f;
      if (zz < ro.z_cut) {
        // w = b*m/(T+eps): dL/db = (dL/dw - D)/(T+eps);  b = sig(a)sig(-a), a = sdf/trunc: db/dsdf = b(1-2 sig(a))/trunc
        const float a = __fdiv_rn(sdf, P.trunc);
        const float sg = sigmoidf_(a);
        const float b = sg * sigmoidf_(-a);
        gs = (dw - D) * inv_den * b * (1.0f - 2.0f * sg) / P.trunc;
      }
      if (zz < lo) {
        gs += g_fsl * fs_w * 2.0f * (sdf - 1.0f) / NS;
      } else if (!(zz > hi) && td > 0.0f) {
        gs += g_sdfl * sdf_w * 2.0f * (__fadd_rn(zz, __fmul_rn(sdf, tr)) - td) * tr / NS;
      }
      g[3] = gs;
      float* o = draw + (ray * S + s) * 5;
#pragma unroll
      for (int k = 0; k < 5; ++k) o[k] = g[k];
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// hash-table scatter: thread = (point, quad of levels), quad fastest: the four threads of a point each take four
// levels (one 32-byte slice of the point's feature gradient), so the point fetch / normalisation is amortised over 32
// reductions per thread and each thread has 32 independent red.global in flight.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) encode_bwd_kernel(const __grid_constant__ DevPlan P, const float2* __restrict__ grid,
                                                         const PointSource src, int64_t n_pts,
                                                         const float4* __restrict__ dfeat, float scale_all,
                                                         float2* __restrict__ dgrid, float* __restrict__ dx) {
  __shared__ DevLevel s_lv[NRT_L];
  if (threadIdx.x < NRT_L) s_lv[threadIdx.x] = P.lv[threadIdx.x];
  __syncthreads();
  const int64_t tt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t pt = tt >> 2;
  const int q = (int)(tt & 3);
  if (pt >= n_pts) return;                  // whole groups of four threads leave together
  const unsigned live = __activemask();
  float x0, x1, x2;
  fetch_point(P, src, pt, x0, x1, x2);
  const float4 ga = __ldg(dfeat + pt * 8 + q * 2), gb = __ldg(dfeat + pt * 8 + q * 2 + 1);
  const float gl[4][2] = {{ga.x * scale_all, ga.y * scale_all}, {ga.z * scale_all, ga.w * scale_all},
                          {gb.x * scale_all, gb.y * scale_all}, {gb.z * scale_all, gb.w * scale_all}};
  float gx[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int li = 0; li < 4; ++li) {
    const DevLevel& L = s_lv[q * 4 + li];
    const float g0 = gl[li][0], g1 = gl[li][1];
    const bool nz = g0 != 0.f || g1 != 0.f;
    if (!nz && !dx) continue;
    uint32_t idx[8];
    float w[8];
    level_corners(L, x0, x1, x2, idx, w);
    if (dgrid && nz) {
      float2* base = dgrid + L.offset;
#pragma unroll
      for (int c = 0; c < 8; c += 2) red_add_xpair(base, idx[c], idx[c + 1], w[c] * g0, w[c] * g1, w[c + 1] * g0, w[c + 1] * g1);
    }
    if (dx) {
      // d out/d x_d = scale * sign_d * prod_{d' != d} w_d' * value
      const LevelPos p = level_pos(L, x0, x1, x2);
      float lx[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float2 v = ldg_f2(grid + L.offset + idx[c]);
        const float dotv = v.x * g0 + v.y * g1;
        const float wx = (c & 1) ? p.f[0] : 1.0f - p.f[0], wy = (c & 2) ? p.f[1] : 1.0f - p.f[1], wz = (c & 4) ? p.f[2] : 1.0f - p.f[2];
        lx[0] += ((c & 1) ? 1.0f : -1.0f) * wy * wz * dotv;
        lx[1] += ((c & 2) ? 1.0f : -1.0f) * wx * wz * dotv;
        lx[2] += ((c & 4) ? 1.0f : -1.0f) * wx * wy * dotv;
      }
#pragma unroll
      for (int d = 0; d < 3; ++d) gx[d] = fmaf(lx[d], L.scale, gx[d]);
    }
  }
  if (dx) {
    // reduce the four level-quads of this point, then one store
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      float v = gx[d];
      v += __shfl_xor_sync(live, v, 1, 4);
      v += __shfl_xor_sync(live, v, 2, 4);
      if (q == 0) dx[pt * 3 + d] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
int launch_composite_bwd(const NrtPlan* plan, const NrtRenderOut* rend, const float* target_rgb, const float* target_d,
                         int64_t n_rays, const double* stats, const float* loss_grad, float* draw, cudaStream_t st) {
  const size_t smem = (size_t)8 * plan->dev.S * 7 * sizeof(float);
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    NRT_CUDA_CHECK(cudaFuncSetAttribute(composite_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  }
  int64_t blocks = (n_rays + 7) / 8;
  int64_t cap = (int64_t)plan->sm_count * 8;
  if (blocks > cap) blocks = cap;
  composite_bwd_kernel<<<(unsigned)blocks, 256, smem, st>>>(plan->dev, *rend, target_rgb, target_d, n_rays, stats, loss_grad, draw);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}

int launch_encode_bwd(const NrtPlan* plan, const float* grid, const PointSource& src, int64_t n_pts, const float* dfeat,
                      float scale_all, float* dgrid, float* dx, cudaStream_t st) {
  if (n_pts == 0) return NRT_OK;
  int64_t threads = n_pts * 4;
  encode_bwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(plan->dev, (const float2*)grid, src, n_pts,
                                                                       (const float4*)dfeat, scale_all, (float2*)dgrid, dx);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}
