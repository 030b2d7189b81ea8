// Tensor-core forward path: point decode (query_sdf / query_color_sdf) and the fused render_rays kernel.
//
// One CTA = 256 threads = one tile of 128 sample points; the thread pair (r, r+128) owns point r = tensor-memory lane r,
// each thread taking half of the hash levels / OneBlob dims / accumulator columns (mlp_tc.cuh).
//   SIMT part (per thread): ray march -> normalise -> 16-level hash gather (+ uncertainty trilerp) -> OneBlob;
//                           the encoded row is split into tf32 hi/lo pieces and staged straight into TMEM (tcgen05.st).
//   tcgen05 part (one elected thread issues): the four bias-free layers as A[tmem] x B[smem] tf32 MMAs with M=128,
//                           three passes each (hi*hi + lo*hi + hi*lo, fp32 accumulate in TMEM) so the result stays within
//                           ~1e-6 of the fp32 reference; ReLU / re-split epilogues read the accumulator back with tcgen05.ld.
//   compositing (render kernel): raw = [rgb logits, sdf, uncertainty] of a block of rays is kept in shared memory and
//                           integrated warp-per-ray with shuffle reductions (sdf2weights + raw2outputs).
//
// TMEM columns (256 allocated per CTA, two CTAs per SM):
//   [0,64)    accumulator of the current phase
//   [64,144)  A_hi: X0[32] (hash features, later relu(h1), later relu(h3)) | OneBlob[48]
//   [144,224) A_lo: same structure, the low-order pieces
// Three dependent tensor-core phases per tile (layer fusion, mlp_tc.cuh): h1 from X0|OneBlob (K=80, N=32); o and a3
// together from h1|OneBlob (K=80, N=48); rgb logits from relu(a3) in X0 (K=32, N=16).
//
// Reference semantics: tp/model/scene_rep.py:160-178 (run_network), src/slam/coslam/model/scene_rep.py:58-64,98-148
// (calc_embedding / query_sdf / query_color_sdf), src/slam/coslam/model/decoder.py:29-41,99-116, and for the ray kernel
// src/slam/coslam/model/scene_rep.py:150-225,66-96 with tp/model/scene_rep.py:64-84.
#include <stdlib.h>

#include "common.cuh"
#include "mlp_tc.cuh"

#define TC_COLS 256               // TMEM columns per CTA (two CTAs per SM)
#define UNIT_PTS_MAX 2048         // sample points of one ray block staged in shared memory

// One tile: every thread of the CTA calls this with the point of its row (both threads of a pair pass the same point;
// inactive rows feed zeros).  Half 0 (threads 0..127) gathers level groups 0 and 2, OneBlob dims 0 and 1, and owns the
// low half of every accumulator (incl. sdf and the rgb logits); half 1 gathers groups 1 and 3, OneBlob dim 2, samples the
// uncertainty grid and owns the high halves.  Outputs land in `out` of the half that computed them.
template <bool COLOR>
__device__ __forceinline__ void decode_tile(const DevPlan& P, TileCtx& c, const float2* __restrict__ grid,
                                            const float* __restrict__ ug, bool active, float x0, float x1, float x2,
                                            float4* __restrict__ feat_out, int64_t feat_pt, PointOut& out,
                                            uint16_t* __restrict__ mask_out = nullptr) {
  const int half = tc_half();
  // ---- encodings -> TMEM ----
  // two hash levels per iteration, gathered pair-cooperatively (common.cuh: gather_levels_paired); rolled loops keep the
  // tile body inside the instruction cache.  Inactive rows sit at x = 0: finite values, results dropped.
#pragma unroll 1
  for (int gi = 0; gi < 4; ++gi) {
    const int g = 2 * (gi >> 1) + half;         // this half's groups of four levels: {half, half + 2}
    const int l0 = 4 * g + 2 * (gi & 1);
    float f[4];
    gather_levels_paired<2>(P.lv + l0, grid, x0, x1, x2, f);
    if (feat_out) feat_out[feat_tiled_index(feat_pt, l0 >> 1)] = make_float4(f[0], f[1], f[2], f[3]);
    stage4(c, TA_X0 + 2 * l0, f);
  }
#pragma unroll 1
  for (int d = 2 * half; d < 2 + half; ++d) {
    float bins[NRT_BINS];
    oneblob16_fast(d == 0 ? x0 : d == 1 ? x1 : x2, bins);
    stage16(c, TA_OB + 16 * d, bins);
  }
  out.unc = (half == 1 && active) ? uncert_sample(P, ug, x0, x1, x2) : 0.f;
  // ---- phase 1: h1 = relu(W1 [hash | oneblob]) ----
  run_layer<80, 32>(c, TA_X0, c.w_hi + FW_W1 * 4, c.w_lo + FW_W1 * 4);
  {
    float h[16];
    tmem_ld16(c.lane_tb + TC_ACC + 16 * half, h);
    tmem_ld_wait();
    uint32_t m = 0u;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      h[j] = fmaxf(h[j], 0.f);
      if (h[j] > 0.f) m |= 1u << j;
    }
    if (mask_out) mask_out[half] = (uint16_t)m;        // NrtRenderOut::masks[pt] = {m1, m3}, this half's 16 bits of m1
    stage16(c, TA_X0 + 16 * half, h);
  }
  // ---- phase 2 on [h1 | oneblob]: o = W2 h1 (columns 0..15) and a3 = W23 h1 + W3_ob oneblob (columns 16..47) ----
  run_layer<80, 48>(c, TA_X0, c.w_hi + FW_W23 * 4, c.w_lo + FW_W23 * 4);
  tmem_ld8(c.lane_tb + TC_ACC + 8 * half, out.o8);        // o[0] = sdf, o[1..15] = geo
  if (COLOR) {
    float h[16];
    tmem_ld16(c.lane_tb + TC_ACC + 16 + 16 * half, h);
    tmem_ld_wait();
    uint32_t m = 0u;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      h[j] = fmaxf(h[j], 0.f);
      if (h[j] > 0.f) m |= 1u << j;
    }
    if (mask_out) mask_out[2 + half] = (uint16_t)m;
    stage16(c, TA_X0 + 16 * half, h);
    // ---- phase 3: rgb logits = W4 relu(a3) ----
    run_layer<32, 16>(c, TA_X0, c.w_hi + FW_W4 * 4, c.w_lo + FW_W4 * 4);
    if (half == 0) {
      float r[4];
      tmem_ld4(c.lane_tb + TC_ACC, r);
      tmem_ld_wait();
      out.rgb[0] = r[0];
      out.rgb[1] = r[1];
      out.rgb[2] = r[2];
    }
  } else {
    tmem_ld_wait();
    out.rgb[0] = out.rgb[1] = out.rgb[2] = 0.f;
  }
  // No barrier needed before the next tile: its tcgen05.st target only A columns whose last reader (this tile's MMAs)
  // has completed, and the accumulator is next written by an MMA issued after the next run_layer() barrier.
}

// ---------------------------------------------------------------------------------------------
// point decode.  Points come from an array x[n,3] (already normalised to the bound) or, when x == NULL, from the regular
// lattice of get_map_volumes (src/slam/coslam/coslam_utils.py:58-97): torch.linspace per axis, meshgrid 'ij', normalised.
// ---------------------------------------------------------------------------------------------
struct LatticeSrc {
  int n[3];           // lattice points per axis
  float lo[3], hi[3]; // float32 bound
  float step[3];      // float32((hi - lo) / (n - 1)), the linspace step
  float* vol_uncert;  // [n0,n1,n2]: softplus(uncert) + 0.01 where 0 <= sdf < 0.5, else 0
  float* vol_sdf;     // [n0,n1,n2]
};

template <bool COLOR>
__global__ void __launch_bounds__(TC_THREADS, 2) points_fwd_tc_kernel(const __grid_constant__ DevPlan P, const NrtParams prm,
                                                                      const float* __restrict__ x, int64_t n,
                                                                      float* __restrict__ raw, float* __restrict__ sdf_uncert,
                                                                      float* __restrict__ geo, const LatticeSrc lat) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  float* rest;
  TileCtx c = cta_prologue<TC_COLS>(smem_raw, prm, &rest);
  cta_prologue_finish(smem_raw, c);
  const float2* grid = reinterpret_cast<const float2*>(prm.grid);
  const int half = tc_half();
  const int64_t n_tiles = (n + 127) / 128;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t pt = tile * 128 + tc_row();
    const bool active = pt < n;
    float x0 = 0.f, x1 = 0.f, x2 = 0.f;
    if (active) {
      if (x) {
        x0 = __ldg(x + pt * 3);
        x1 = __ldg(x + pt * 3 + 1);
        x2 = __ldg(x + pt * 3 + 2);
      } else {
        const int k2 = (int)(pt % lat.n[2]), k1 = (int)((pt / lat.n[2]) % lat.n[1]), k0 = (int)(pt / ((int64_t)lat.n[2] * lat.n[1]));
        x0 = normalise1(P, 0, linspace_at(lat.lo[0], lat.hi[0], lat.step[0], lat.n[0], k0));
        x1 = normalise1(P, 1, linspace_at(lat.lo[1], lat.hi[1], lat.step[1], lat.n[1], k1));
        x2 = normalise1(P, 2, linspace_at(lat.lo[2], lat.hi[2], lat.step[2], lat.n[2], k2));
      }
    }
    PointOut o;
    decode_tile<COLOR>(P, c, grid, prm.uncert, active, x0, x1, x2, nullptr, 0, o);
    if (active && lat.vol_sdf && half == 0) {
      // get_map_volumes: uncertainty only where the point is just outside the surface
      const float sdf = o.o8[0];
      const float u = softplusf_(uncert_sample(P, prm.uncert, x0, x1, x2)) + 0.01f;
      lat.vol_sdf[pt] = sdf;
      lat.vol_uncert[pt] = (sdf >= 0.0f && sdf < 0.5f) ? u : 0.0f;
    }
    if (active) {
      if (half == 0) {
        if (raw) {
          float* r = raw + pt * 5;
          r[0] = o.rgb[0];
          r[1] = o.rgb[1];
          r[2] = o.rgb[2];
          r[3] = o.o8[0];
        }
        if (sdf_uncert) sdf_uncert[pt * 2] = o.o8[0];
        if (geo) {
#pragma unroll
          for (int k = 0; k < 7; ++k) geo[pt * NRT_GEO + k] = o.o8[1 + k];
        }
      } else {
        if (raw) raw[pt * 5 + 4] = o.unc;
        if (sdf_uncert) sdf_uncert[pt * 2 + 1] = o.unc;
        if (geo) {
#pragma unroll
          for (int k = 0; k < 8; ++k) geo[pt * NRT_GEO + 7 + k] = o.o8[k];
        }
      }
    }
  }
  cta_epilogue<TC_COLS>(c);
}

// ---------------------------------------------------------------------------------------------
// fused render_rays: a CTA takes blocks of `rpu` consecutive rays (<= UNIT_PTS_MAX sample points);
// smem after the weights: [ray o,d: rpu*6 | z: rpu*S | raw: rpu*S*5]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 2) render_fwd_tc_kernel(const __grid_constant__ DevPlan P, const NrtParams prm,
                                                                      const float* __restrict__ rays_o,
                                                                      const float* __restrict__ rays_d,
                                                                      const float* __restrict__ target_d, int64_t n_rays,
                                                                      const float* __restrict__ z_in, const float* __restrict__ u,
                                                                      int perturb, uint64_t seed, const int* __restrict__ seed_step, int rpu,
                                                                      const NrtRenderOut out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  float* rest;
  TileCtx c = cta_prologue<TC_COLS>(smem_raw, prm, &rest);
  cta_prologue_finish(smem_raw, c);
  if (seed_step) seed ^= (uint64_t)(uint32_t)__ldg(seed_step) * 0x9E3779B97F4A7C15ull;
  const int S = P.S;
  float* s_ray = rest;
  float* s_z = s_ray + rpu * 6;
  float* s_raw = s_z + rpu * S;
  const float2* grid = reinterpret_cast<const float2*>(prm.grid);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, half = tc_half();
  constexpr int kWarps = TC_THREADS / 32;
  const int64_t n_units = (n_rays + rpu - 1) / rpu;

  for (int64_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    const int64_t r0 = unit * rpu;
    const int nr = (int)min((int64_t)rpu, n_rays - r0);
    const int npts = nr * S;
    // ---- stage the rays and their depth samples ----
    for (int i = threadIdx.x; i < nr * 6; i += TC_THREADS) {
      const int rl = i / 6, k = i - rl * 6;
      s_ray[i] = k < 3 ? __ldg(rays_o + (r0 + rl) * 3 + k) : __ldg(rays_d + (r0 + rl) * 3 + k - 3);
    }
    for (int rl = warp; rl < nr; rl += kWarps) {
      const int64_t ray = r0 + rl;
      float* z = s_z + rl * S;
      if (z_in) {
        for (int s = lane; s < S; s += 32) z[s] = __ldg(z_in + ray * S + s);
      } else {
        warp_sample_z(P, __ldg(target_d + ray), u ? u + ray * S : nullptr, perturb, seed, ray, z, lane);
      }
    }
    __syncthreads();
    // ---- decode tile by tile ----
    for (int t0 = 0; t0 < npts; t0 += 128) {
      const int pl = t0 + tc_row();
      const bool active = pl < npts;
      float x0 = 0.f, x1 = 0.f, x2 = 0.f;
      if (active) {
        const int rl = pl / S;
        const float zz = s_z[pl];
        const float* ry = s_ray + rl * 6;
        // pts = o + d*z then (pts - bb_min)/(bb_max - bb_min): separate roundings, like the reference's tensor ops
        x0 = normalise1(P, 0, __fadd_rn(ry[0], __fmul_rn(ry[3], zz)));
        x1 = normalise1(P, 1, __fadd_rn(ry[1], __fmul_rn(ry[4], zz)));
        x2 = normalise1(P, 2, __fadd_rn(ry[2], __fmul_rn(ry[5], zz)));
      }
      PointOut po;
      decode_tile<true>(P, c, grid, prm.uncert, active, x0, x1, x2,
                        (out.feat && active) ? reinterpret_cast<float4*>(out.feat) : nullptr, r0 * S + pl, po,
                        (out.masks && active) ? reinterpret_cast<uint16_t*>(out.masks) + (r0 * S + pl) * 4 : nullptr);
      if (active) {
        float* r = s_raw + pl * 5;
        if (half == 0) {
          r[0] = po.rgb[0];
          r[1] = po.rgb[1];
          r[2] = po.rgb[2];
          r[3] = po.o8[0];
        } else {
          r[4] = po.unc;
        }
      }
    }
    __syncthreads();
    // ---- integrate along each ray ----
    for (int rl = warp; rl < nr; rl += kWarps) {
      const int64_t ray = r0 + rl;
      const float* z = s_z + rl * S;
      const float* raw = s_raw + rl * S * 5;
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
      if (out.z_vals)
        for (int s = lane; s < S; s += 32) out.z_vals[ray * S + s] = z[s];
      if (out.raw)
        for (int i = lane; i < S * 5; i += 32) out.raw[ray * S * 5 + i] = raw[i];
    }
    __syncthreads();
  }
  cta_epilogue<TC_COLS>(c);
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
// Both kernels are held to two CTAs per SM (2 x 256 TMEM columns = the whole tensor memory): the point kernel
// requests enough dynamic shared memory that a third CTA cannot become resident and then stall in tcgen05.alloc.
static const size_t kPointsSmem = 80 * 1024;

int launch_decode_fwd(const NrtPlan* plan, const NrtParams* prm, const float* x, int64_t n, int with_color, float* raw,
                      float* sdf_uncert, float* geo, cudaStream_t st) {
  if (n == 0) return NRT_OK;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    NRT_CUDA_CHECK(cudaFuncSetAttribute(points_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPointsSmem));
    NRT_CUDA_CHECK(cudaFuncSetAttribute(points_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPointsSmem));
  }
  const int64_t tiles = (n + 127) / 128;
  const int blocks = (int)(tiles < 2 * plan->sm_count ? tiles : 2 * plan->sm_count);
  LatticeSrc none{};
  if (with_color)
    points_fwd_tc_kernel<true><<<blocks, TC_THREADS, kPointsSmem, st>>>(plan->dev, *prm, x, n, raw, sdf_uncert, geo, none);
  else
    points_fwd_tc_kernel<false><<<blocks, TC_THREADS, kPointsSmem, st>>>(plan->dev, *prm, x, n, raw, sdf_uncert, geo, none);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}

// dense uncertainty + SDF sweep over the lattice of get_map_volumes; dims = lattice points per axis
int launch_map_volumes(const NrtPlan* plan, const NrtParams* prm, const int* dims, float* vol_uncert, float* vol_sdf,
                       cudaStream_t st) {
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    NRT_CUDA_CHECK(cudaFuncSetAttribute(points_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPointsSmem));
  }
  LatticeSrc lat{};
  for (int a = 0; a < 3; ++a) {
    lat.n[a] = dims[a];
    lat.lo[a] = plan->cfg.bound_min[a];
    lat.hi[a] = plan->cfg.bound_max[a];
    lat.step[a] = dims[a] > 1 ? (lat.hi[a] - lat.lo[a]) / (float)(dims[a] - 1) : 0.f;
  }
  lat.vol_uncert = vol_uncert;
  lat.vol_sdf = vol_sdf;
  const int64_t n = (int64_t)dims[0] * dims[1] * dims[2];
  const int64_t tiles = (n + 127) / 128;
  const int blocks = (int)(tiles < 2 * plan->sm_count ? tiles : 2 * plan->sm_count);
  points_fwd_tc_kernel<false><<<blocks, TC_THREADS, kPointsSmem, st>>>(plan->dev, *prm, nullptr, n, nullptr, nullptr, nullptr, lat);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}

int launch_render_fwd_ws(const NrtPlan*, const NrtParams*, const float*, const float*, const float*, int64_t, const float*,
                         const float*, int, uint64_t, const NrtRenderOut*, const float*, double*, float*, const int*, cudaStream_t);   // forward_ws.cu
int launch_loss_partial(const NrtPlan*, const NrtRenderOut*, const float*, const float*, int64_t, double*, cudaStream_t);
int launch_loss_finalize(const double*, float*, cudaStream_t);

// target_rgb / stats non-NULL: also produce the loss statistics of the shard (nrt_render_fwd_stats)
int launch_render_fwd(const NrtPlan* plan, const NrtParams* prm, const float* rays_o, const float* rays_d,
                      const float* target_d, int64_t n_rays, const float* z_in, const float* u, int perturb, uint64_t seed,
                      const NrtRenderOut* out, const float* target_rgb, double* stats, float* losses, const int* seed_step,
                      cudaStream_t st) {
  if (n_rays == 0) return NRT_OK;
  // The warp-specialised kernel (forward_ws.cu) is the product path; NRT_RENDER_IMPL=tc selects the one-role kernel below
  // (same arithmetic, bit-identical results) for A/B measurements.
  static int use_ws = -1;
  if (use_ws < 0) {
    const char* e = getenv("NRT_RENDER_IMPL");
    use_ws = (e && e[0] == 't' && e[1] == 'c' && e[2] == 0) ? 0 : 1;
  }
  if (use_ws) return launch_render_fwd_ws(plan, prm, rays_o, rays_d, target_d, n_rays, z_in, u, perturb, seed, out, target_rgb, stats, losses, seed_step, st);
  if (target_rgb) {
    if (int rc = launch_render_fwd(plan, prm, rays_o, rays_d, target_d, n_rays, z_in, u, perturb, seed, out, nullptr, nullptr, nullptr, seed_step, st)) return rc;
    if (int rc = launch_loss_partial(plan, out, target_rgb, target_d, n_rays, stats, st)) return rc;
    return losses ? launch_loss_finalize(stats, losses, st) : NRT_OK;
  }
  const int S = plan->dev.S;
  const int slots = 2 * plan->sm_count;
  // rays per block: as many blocks as there are CTA slots (one balanced wave), capped by the staging buffer
  int64_t rpu = (n_rays + slots - 1) / slots;
  const int cap = UNIT_PTS_MAX / S > 1 ? UNIT_PTS_MAX / S : 1;
  if (rpu > cap) rpu = cap;
  if (rpu < 1) rpu = 1;
  const int64_t units = (n_rays + rpu - 1) / rpu;
  // Shared memory per CTA also decides the L1 size: two CTAs x (this + 1 KB) must stay inside the 196 KB carve-out, which
  // leaves 32 KB of L1 for the gather.  Measured (B200, 4096 x 128): 0.207 ms with 32 KB or 64 KB of L1, 0.251 ms when the
  // carve-out is forced to the full 228 KB; and any static shared memory in this kernel tips it over that edge.
  size_t smem = TC_SMEM_WEIGHTS + (size_t)rpu * (6 + 6 * S) * sizeof(float);
  if (smem < kPointsSmem) smem = kPointsSmem;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    NRT_CUDA_CHECK(cudaFuncSetAttribute(render_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
  }
  const int blocks = (int)(units < slots ? units : slots);
  render_fwd_tc_kernel<<<blocks, TC_THREADS, smem, st>>>(plan->dev, *prm, rays_o, rays_d, target_d, n_rays, z_in, u, perturb, seed,
                                                  seed_step, (int)rpu, *out);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}
