// Backward kernels of the mapping iteration:
//   composite_bwd_kernel : dL/d raw[B,S,5] from the loss definition (per ray, one warp)
//   decode_bwd_kernel    : recompute MLP activations from saved hash features, back-propagate through both
//                          MLPs, accumulate dW1..dW4 with a shared-memory tile GEMM, emit dL/d hash-features
//                          and scatter the uncertainty-grid gradient
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

  for (int64_t ray = (int64_t)blockIdx.x * wpb + warp; ray < n_rays; ray += (int64_t)gridDim.x * wpb) {
    for (int s = lane; s < S; s += 32) z[s] = rend.z_vals[ray * S + s];
    for (int i = lane; i < S * 5; i += 32) raw[i] = rend.raw[ray * S * 5 + i];
    __syncwarp();
    RayOut ro = warp_composite(P, S, raw, z, wbuf, lane);
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
    for (int s = lane; s < S; s += 32) {
      const float w = wbuf[s];
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
      const float unc = softplusf_(ur) + 0.01f;
      float c[3], g[5];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        c[k] = sigmoidf_(raw[s * 5 + k]);
        g[k] = g_rgb[k] * w * c[k] * (1.0f - c[k]);
      }
      float dw = g_rgb[0] * c[0] + g_rgb[1] * c[1] + g_rgb[2] * c[2] + g_depth * zz + g_U * 2.0f * w * unc;
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
      g[4] = g_U * w * w * (ur > 20.0f ? 1.0f : sigmoidf_(ur));
      float* o = draw + (ray * S + s) * 5;
#pragma unroll
      for (int k = 0; k < 5; ++k) o[k] = g[k];
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// MLP backward over tiles of 128 points (thread = point in the per-point phase, then all threads share
// the four weight-gradient tile GEMMs).  smem rows are [feature][point] with pitch TP.
// ---------------------------------------------------------------------------------------------
#define TILE 128
#define TP 132
#define R_IN1 0                 // 80 rows: hash 0..31 | oneblob 32..79
#define R_H1 (R_IN1 + 80)       // 32
#define R_GEO (R_H1 + 32)       // 16 (row 15 = 0)
#define R_H3 (R_GEO + 16)       // 32
#define R_DA1 (R_H3 + 32)       // 32
#define R_DO (R_DA1 + 32)       // 16
#define R_DA3 (R_DO + 16)       // 32
#define R_DC (R_DA3 + 32)       // 4 (row 3 = 0)
#define R_TOTAL (R_DC + 4)

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float dot4(float4 a, float4 b, float acc) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  return fmaf(a.w, b.w, acc);
}

__global__ void __launch_bounds__(TILE, 1) decode_bwd_kernel(const __grid_constant__ DevPlan P, const NrtParams prm,
                                                             const PointSource src, int64_t n_pts,
                                                             const float* feat, const float* __restrict__ draw,
                                                             float* dfeat, const NrtGrads grads) {
  extern __shared__ __align__(16) float smem[];
  float* sw = smem;
  float* tile = smem + SW_BWD_FLOATS;
  load_weights_smem(sw, prm, true);
  const int t = threadIdx.x;
  // weight-gradient ownership (see the bank analysis in DESIGN.md): strided feature assignment keeps the
  // eight lanes of a quarter-warp on distinct bank groups
  const int jb = t >> 4, kb = t & 15;          // dW1/dW3: j in [4jb,4jb+4), k = kb + 16*kk
  const int i2 = t >> 3, jq = t & 7;           // dW2: i = i2, j = jq + 8*jj
  const int i4 = t >> 5, j4 = t & 31;          // dW4: (i4, j4)
  float acc1[4][5], acc3[4][4], acc2[4], acc4 = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
#pragma unroll
    for (int b = 0; b < 5; ++b) acc1[a][b] = 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) acc3[a][b] = 0.f;
    acc2[a] = 0.f;
  }
  __syncthreads();

  const int64_t n_tiles = (n_pts + TILE - 1) / TILE;
  for (int64_t tl = blockIdx.x; tl < n_tiles; tl += gridDim.x) {
    const int64_t pt = tl * TILE + t;
    const bool active = pt < n_pts;
    // ---------------- per-point phase ----------------
    {
      float x0 = 0.f, x1 = 0.f, x2 = 0.f;
      if (active) fetch_point(P, src, pt, x0, x1, x2);
      float h[NRT_H], a3[NRT_H];
#pragma unroll
      for (int j = 0; j < NRT_H; ++j) {
        h[j] = 0.f;
        a3[j] = 0.f;
      }
      // saved hash features -> tile + layer 1
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 f = active ? __ldg(reinterpret_cast<const float4*>(feat + pt * NRT_ENC) + q) : make_float4(0, 0, 0, 0);
        float fv[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          tile[(R_IN1 + 4 * q + e) * TP + t] = fv[e];
          fma_row32(fv[e], sw + SW_W1T + (4 * q + e) * 32, h);
        }
      }
      float xs[3] = {x0, x1, x2};
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        float bins[NRT_BINS];
        oneblob16(xs[d], bins);
#pragma unroll
        for (int b = 0; b < NRT_BINS; ++b) {
          const float v = active ? bins[b] : 0.f;
          tile[(R_IN1 + NRT_ENC + d * NRT_BINS + b) * TP + t] = v;
          fma_row32(v, sw + SW_W1T + (NRT_ENC + d * NRT_BINS + b) * 32, h);
          fma_row32(v, sw + SW_W3T + (d * NRT_BINS + b) * 32, a3);
        }
      }
      float o[NRT_O];
#pragma unroll
      for (int i = 0; i < NRT_O; ++i) o[i] = 0.f;
      unsigned m1 = 0u;     // ReLU masks
#pragma unroll
      for (int j = 0; j < NRT_H; ++j) {
        const float v = fmaxf(h[j], 0.f);
        if (v > 0.f) m1 |= 1u << j;
        tile[(R_H1 + j) * TP + t] = v;
        const float4* r4 = reinterpret_cast<const float4*>(sw + SW_W2T + j * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 w = r4[q];
          o[4 * q + 0] = fmaf(v, w.x, o[4 * q + 0]);
          o[4 * q + 1] = fmaf(v, w.y, o[4 * q + 1]);
          o[4 * q + 2] = fmaf(v, w.z, o[4 * q + 2]);
          o[4 * q + 3] = fmaf(v, w.w, o[4 * q + 3]);
        }
      }
#pragma unroll
      for (int k = 0; k < NRT_GEO; ++k) {
        tile[(R_GEO + k) * TP + t] = o[1 + k];
        fma_row32(o[1 + k], sw + SW_W3T + (NRT_OB + k) * 32, a3);
      }
      tile[(R_GEO + 15) * TP + t] = 0.f;
      // upstream gradient of this point
      float dc[3] = {0.f, 0.f, 0.f}, dsdf = 0.f, du = 0.f;
      if (active) {
        const float* g = draw + pt * 5;
        dc[0] = __ldg(g);
        dc[1] = __ldg(g + 1);
        dc[2] = __ldg(g + 2);
        dsdf = __ldg(g + 3);
        du = __ldg(g + 4);
      }
      tile[(R_DC + 0) * TP + t] = dc[0];
      tile[(R_DC + 1) * TP + t] = dc[1];
      tile[(R_DC + 2) * TP + t] = dc[2];
      tile[(R_DC + 3) * TP + t] = 0.f;
      // colour net backward: dh3 = W4^T dc, da3 = dh3 * relu'
      float da3[NRT_H];
#pragma unroll
      for (int j = 0; j < NRT_H; ++j) {
        const float hv = fmaxf(a3[j], 0.f);
        tile[(R_H3 + j) * TP + t] = hv;
        float d = dc[0] * sw[SW_W4 + j] + dc[1] * sw[SW_W4 + 32 + j] + dc[2] * sw[SW_W4 + 64 + j];
        da3[j] = hv > 0.f ? d : 0.f;
        tile[(R_DA3 + j) * TP + t] = da3[j];
      }
      // d geo = W3[:,48:63]^T da3 ; do = [dsdf, dgeo]
      float dov[NRT_O];
      dov[0] = dsdf;
#pragma unroll
      for (int k = 0; k < NRT_GEO; ++k) {
        const float4* r4 = reinterpret_cast<const float4*>(sw + SW_W3T + (NRT_OB + k) * 32);
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 w = r4[q];
          s = fmaf(w.x, da3[4 * q], s);
          s = fmaf(w.y, da3[4 * q + 1], s);
          s = fmaf(w.z, da3[4 * q + 2], s);
          s = fmaf(w.w, da3[4 * q + 3], s);
        }
        dov[1 + k] = s;
      }
#pragma unroll
      for (int i = 0; i < NRT_O; ++i) tile[(R_DO + i) * TP + t] = dov[i];
      // SDF net backward: dh1 = W2^T do, da1 = dh1 * relu'   (a3[] is dead: reuse the registers as da1)
      float* da1 = a3;
#pragma unroll
      for (int j = 0; j < NRT_H; ++j) da1[j] = 0.f;
#pragma unroll
      for (int i = 0; i < NRT_O; ++i) fma_row32(dov[i], sw + SW_W2 + i * 32, da1);
#pragma unroll
      for (int j = 0; j < NRT_H; ++j) {
        da1[j] = (m1 >> j) & 1u ? da1[j] : 0.f;
        tile[(R_DA1 + j) * TP + t] = da1[j];
      }
      // d hash features = W1[:, :32]^T da1
      if (active) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float r[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float4* r4 = reinterpret_cast<const float4*>(sw + SW_W1T + (4 * q + e) * 32);
            float s = 0.f;
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) {
              float4 w = r4[qq];
              s = fmaf(w.x, da1[4 * qq], s);
              s = fmaf(w.y, da1[4 * qq + 1], s);
              s = fmaf(w.z, da1[4 * qq + 2], s);
              s = fmaf(w.w, da1[4 * qq + 3], s);
            }
            r[e] = s;
          }
          if (dfeat) reinterpret_cast<float4*>(dfeat + pt * NRT_ENC)[q] = make_float4(r[0], r[1], r[2], r[3]);
        }
        // uncertainty grid: raw[...,4] is the trilinear sample itself
        if (grads.uncert && du != 0.f) {
          UncertPos up = uncert_pos(P, x0, x1, x2);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float w;
            int off = uncert_corner(P, up, c, w);
            if (off >= 0) atomicAdd(grads.uncert + off, w * du);
          }
        }
      }
    }
    __syncthreads();
    // ---------------- weight-gradient tile GEMMs: dW[j][k] += sum_r dA[j][r] * IN[k][r] ----------------
#pragma unroll 2
    for (int r = 0; r < TILE; r += 4) {
      float4 a[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) a[jj] = lds4(tile + (R_DA1 + 4 * jb + jj) * TP + r);
#pragma unroll
      for (int kk = 0; kk < 5; ++kk) {
        float4 b = lds4(tile + (R_IN1 + kb + 16 * kk) * TP + r);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc1[jj][kk] = dot4(a[jj], b, acc1[jj][kk]);
      }
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) a[jj] = lds4(tile + (R_DA3 + 4 * jb + jj) * TP + r);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        // colour-net input k = kb + 16kk: oneblob rows live behind the hash rows of IN1, geo rows in GEO
        float4 b = kk < 3 ? lds4(tile + (R_IN1 + NRT_ENC + kb + 16 * kk) * TP + r) : lds4(tile + (R_GEO + kb) * TP + r);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc3[jj][kk] = dot4(a[jj], b, acc3[jj][kk]);
      }
      {
        float4 d = lds4(tile + (R_DO + i2) * TP + r);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc2[jj] = dot4(d, lds4(tile + (R_H1 + jq + 8 * jj) * TP + r), acc2[jj]);
        acc4 = dot4(lds4(tile + (R_DC + i4) * TP + r), lds4(tile + (R_H3 + j4) * TP + r), acc4);
      }
    }
    __syncthreads();
  }
  // ---------------- flush the per-CTA weight gradients ----------------
  if (grads.w1) {
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
      for (int kk = 0; kk < 5; ++kk) atomicAdd(grads.w1 + (4 * jb + jj) * 80 + kb + 16 * kk, acc1[jj][kk]);
  }
  if (grads.w3) {
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int k = kb + 16 * kk;
        if (k < 63) atomicAdd(grads.w3 + (4 * jb + jj) * 63 + k, acc3[jj][kk]);
      }
  }
  if (grads.w2) {
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) atomicAdd(grads.w2 + i2 * 32 + jq + 8 * jj, acc2[jj]);
  }
  if (grads.w4 && i4 < 3) atomicAdd(grads.w4 + i4 * 32 + j4, acc4);
}

// ---------------------------------------------------------------------------------------------
// hash-table scatter: thread = (point, level), level fastest
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) encode_bwd_kernel(const __grid_constant__ DevPlan P, const float2* __restrict__ grid,
                                                         const PointSource src, int64_t n_pts,
                                                         const float2* __restrict__ dfeat, float scale_all,
                                                         float2* __restrict__ dgrid, float* __restrict__ dx) {
  __shared__ DevLevel s_lv[NRT_L];
  if (threadIdx.x < NRT_L) s_lv[threadIdx.x] = P.lv[threadIdx.x];
  __syncthreads();
  int64_t tt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t pt = tt >> 4;
  int l = (int)(tt & 15);
  if (pt >= n_pts) return;
  const unsigned live = __activemask();     // whole half-warps leave together (16 threads per point)
  float x0, x1, x2;
  fetch_point(P, src, pt, x0, x1, x2);
  float2 g = __ldg(dfeat + pt * NRT_L + l);
  g.x *= scale_all;
  g.y *= scale_all;
  const DevLevel& L = s_lv[l];
  LevelPos p = level_pos(L, x0, x1, x2);
  float2* base = dgrid ? dgrid + L.offset : nullptr;
  float gx[3] = {0.f, 0.f, 0.f};
  const bool nz = g.x != 0.f || g.y != 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint32_t idx = level_index(L, p.g[0] + (c & 1), p.g[1] + ((c >> 1) & 1), p.g[2] + ((c >> 2) & 1));
    if (base && nz) {
      float w = corner_weight(p, c);
      red_add_f2(base + idx, w * g.x, w * g.y);
    }
    if (dx) {
      // d out/d x_d = scale * sign_d * prod_{d' != d} w_d' * value
      float2 v = ldg_f2(grid + L.offset + idx);
      float dotv = v.x * g.x + v.y * g.y;
      float wx = (c & 1) ? p.f[0] : 1.0f - p.f[0], wy = (c & 2) ? p.f[1] : 1.0f - p.f[1], wz = (c & 4) ? p.f[2] : 1.0f - p.f[2];
      gx[0] += ((c & 1) ? 1.0f : -1.0f) * wy * wz * dotv;
      gx[1] += ((c & 2) ? 1.0f : -1.0f) * wx * wz * dotv;
      gx[2] += ((c & 4) ? 1.0f : -1.0f) * wx * wy * dotv;
    }
  }
  if (dx) {
    // reduce the 16 levels of this point inside the half-warp, then one store
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      float v = gx[d] * L.scale;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(live, v, o, 16);
      if (l == 0) dx[pt * 3 + d] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
size_t decode_bwd_smem() { return (size_t)(SW_BWD_FLOATS + R_TOTAL * TP) * sizeof(float); }

int launch_composite_bwd(const NrtPlan* plan, const NrtRenderOut* rend, const float* target_rgb, const float* target_d,
                         int64_t n_rays, const double* stats, const float* loss_grad, float* draw, cudaStream_t st) {
  const size_t smem = (size_t)8 * plan->dev.S * 7 * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    NRT_CUDA_CHECK(cudaFuncSetAttribute(composite_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr_set = true;
  }
  int64_t blocks = (n_rays + 7) / 8;
  int64_t cap = (int64_t)plan->sm_count * 8;
  if (blocks > cap) blocks = cap;
  composite_bwd_kernel<<<(unsigned)blocks, 256, smem, st>>>(plan->dev, *rend, target_rgb, target_d, n_rays, stats, loss_grad, draw);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}

int launch_decode_bwd(const NrtPlan* plan, const NrtParams* prm, const PointSource& src, int64_t n_pts, const float* feat,
                      const float* draw, float* dfeat, const NrtGrads* grads, cudaStream_t st) {
  const size_t smem = decode_bwd_smem();
  static bool attr_set = false;
  if (!attr_set) {
    NRT_CUDA_CHECK(cudaFuncSetAttribute(decode_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  int64_t tiles = (n_pts + TILE - 1) / TILE;
  int64_t blocks = tiles < plan->sm_count ? tiles : plan->sm_count;
  decode_bwd_kernel<<<(unsigned)blocks, TILE, smem, st>>>(plan->dev, *prm, src, n_pts, feat, draw, dfeat, *grads);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}

int launch_encode_bwd(const NrtPlan* plan, const float* grid, const PointSource& src, int64_t n_pts, const float* dfeat,
                      float scale_all, float* dgrid, float* dx, cudaStream_t st) {
  if (n_pts == 0) return NRT_OK;
  int64_t threads = n_pts * NRT_L;
  encode_bwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(plan->dev, (const float2*)grid, src, n_pts,
                                                                       (const float2*)dfeat, scale_all, (float2*)dgrid, dx);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}
