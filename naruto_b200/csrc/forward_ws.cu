// Warp-specialised fused render_rays kernel (reference: src/slam/coslam/model/scene_rep.py:150-225,66-96 with
// tp/model/scene_rep.py:64-84,160-178): depth sampling -> hash gather -> MLPs on the tensor cores -> compositing.
//
// What bounds the fused forward is the L1 gather pipe (l1tex data-pipe wavefronts), and in the one-role kernel
// (forward_tc.cu: render_fwd_tc_kernel) that pipe idles whenever a CTA's warps sit in the three dependent tensor-core phases
// of a tile.  Here the two jobs run on different warps of one persistent CTA per SM, connected by shared-memory rings:
//
//   warps  0.. 3   MLP group 0      one thread = one row (sample point) = one TMEM lane, columns [0, 256) of tensor memory
//   warps  4.. 7   MLP group 1      columns [256, 512)
//   warps  8..15   gather set 0     feeds group 0: warp j -> rows 32 (j%4).., levels 8 (j/4)..+7, pair-cooperative gathers
//   warps 16..23   gather set 1     feeds group 1
//
// A "sub-CTA" = one MLP group + its gather set (384 threads) works through blocks of `rpu` consecutive rays on its own:
// all 12 warps stage the rays and sample the depths (warp per ray); then, tile by tile (128 points), the gather warps write
// the 32 hash features of every row into a ring stage [8 chunks][128 rows][4] (chunk-major, so both the 16-byte stores of
// the gather threads and the 16-byte loads of the row owners are conflict-free) while the MLP group drains the previous
// stage: features + OneBlob -> tf32 hi/lo -> TMEM, three tcgen05 phases (mlp_tc.cuh), raw -> shared memory; finally all 12
// warps integrate along the rays (warp per ray).  Named barriers carry the ring (full / empty per stage), the MLP group's
// publish step and the sub-CTA phases; the two sub-CTAs share only the weights in shared memory and the TMEM allocation.
// The gather warps never wait for a tensor-core phase, so the L1 pipe stays busy; the weights are staged once per SM.
// Measured and dropped for tables beyond L2 (T = 2^21, round 2): `prefetch.global.L2` of the next batch of levels by the gather
// warps (forward 0.83 -> 0.92 ms: the prefetches cost LSU slots the loads need; at T = 2^16 0.173 -> 0.238 ms) and interleaving
// dense and hashed levels between the two halves of a gather set (0.81 -> 0.83 ms).
#include <cstdlib>

#include "common.cuh"
#include "mlp_tc.cuh"

#define WS_THREADS 768
#define WS_SUB 384                  // threads of a sub-CTA: 128 MLP + 256 gather
#define WS_ROWS 128
#define WS_NST 2                    // ring stages per sub-CTA
#define WS_STAGE_FLOATS (8 * 128 * 4)
#define WS_UNIT_PTS 1024            // sample points of one ray block staged in shared memory (per sub-CTA)
#define WS_COLS 512
// named barriers (id 0 = __syncthreads)
#define WS_BAR_MLP(g) (1 + (g))                   // 128 MLP threads of group g
#define WS_BAR_FULL(g, s) (3 + 2 * (g) + (s))     // 256 gather arrive + 128 MLP sync
#define WS_BAR_EMPTY(g, s) (7 + 2 * (g) + (s))    // 128 MLP arrive + 256 gather sync
#define WS_BAR_UNIT(g) (11 + (g))                 // all 384 threads of sub-CTA g

// Optional fused loss statistics (what nrt_loss_partial computes from the materialised outputs; forward.cu:
// loss_partial_kernel): the compositing warp already holds the ray's z / raw in shared memory and its integrals in
// registers.  Lane k of every warp keeps the fp64 partial sum of statistic k in a register (no shared memory while the kernel
// runs: the shared-memory footprint decides the L1 carve-out, and 2 KB more cost the forward 3-6 %); the partials are folded
// per CTA at the end through the then idle ring area and reduced in CTA order by the last CTA to finish, so the result does
// not depend on scheduling.
// profiling aid (NRT_FWD_DEBUG=1): per-CTA cycle accounting of sub-CTA 0, read back with nrt_debug_read(which = 1)
__device__ long long g_ws_trace[256 * 8];

struct LossFuse {
  int trace;                 // profiling aid on
  const float* target_rgb;   // NULL: statistics off
  const float* target_d;
  double* part;              // [gridDim.x][NRT_N_STATS]
  unsigned int* counter;
  double* stats;             // [NRT_N_STATS]
  float* losses;             // optional [NRT_N_LOSS]: single-shard callers get the finalized losses from the last CTA
};
#define WS_STAT_SLOTS 12     // 11 sums + the minimum of uncert_map

template <int K, int N>
__device__ __forceinline__ void ws_run_layer(TileCtx& c, int g, bool issuer, int a_col, uint32_t wh, uint32_t wl) {
  tmem_st_wait();
  tc_fence_before();
  bar_sync(WS_BAR_MLP(g), WS_ROWS);
  if (issuer) {                             // the group's first warp stays converged; one elected lane issues
    tc_fence_after();
    if (elect_one()) {
      issue_layer<K, N>(c, a_col, wh, wl);
      mma_commit(c.bar);
    }
  }
  layer_wait(c);
}

// point of row `pl` of the current ray block (inactive rows sit at x = 0: finite values, results dropped)
__device__ __forceinline__ void ws_point(const DevPlan& P, const float* __restrict__ s_ray, const float* __restrict__ s_z, int pl,
                                         int npts, int S, float& x0, float& x1, float& x2) {
  x0 = x1 = x2 = 0.f;
  if (pl < npts) {
    const int rl = pl / S;
    const float zz = s_z[pl];
    const float* ry = s_ray + rl * 6;
    // pts = o + d*z then (pts - bb_min)/(bb_max - bb_min): separate roundings, like the reference's tensor ops
    x0 = normalise1(P, 0, __fadd_rn(ry[0], __fmul_rn(ry[3], zz)));
    x1 = normalise1(P, 1, __fadd_rn(ry[1], __fmul_rn(ry[4], zz)));
    x2 = normalise1(P, 2, __fadd_rn(ry[2], __fmul_rn(ry[5], zz)));
  }
}

// TRAIN: the launch saves what the backward pass needs (hash features, ReLU masks) and / or accumulates the loss statistics.
// Forward-only launches (evaluation, the uncertainty sweep) run the instantiation without any of that code.
template <bool TRAIN>
__global__ void __launch_bounds__(WS_THREADS, 1) render_fwd_ws_kernel(const __grid_constant__ DevPlan P, const NrtParams prm,
                                                                      const float* __restrict__ rays_o,
                                                                      const float* __restrict__ rays_d,
                                                                      const float* __restrict__ target_d, int64_t n_rays,
                                                                      const float* __restrict__ z_in, const float* __restrict__ u,
                                                                      int perturb, uint64_t seed, const int* __restrict__ seed_step, int rpu,
                                                                      const NrtRenderOut out, const LossFuse lf) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);           // [2] MMA-complete mbarriers, one per MLP group
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem_raw + 16);
  float* sw = reinterpret_cast<float*>(smem_raw + TC_SMEM_HEADER);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const bool tr = lf.trace && blockIdx.x < 256;
  const long long tr_t0 = tr ? clock64() : 0;
  long long tr_stage = 0, tr_tiles = 0, tr_comp = 0, tr_pro = 0;
  if (seed_step) seed ^= (uint64_t)(uint32_t)__ldg(seed_step) * 0x9E3779B97F4A7C15ull;     // per-iteration jitter under graph replay
  if (warp == 0) {
    if (threadIdx.x == 0) {
      mbar_init(&bars[0], 1);
      mbar_init(&bars[1], 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<WS_COLS>(slot);
  }
  // (the ring area behind the weight block is free until the first tile: staging scratch; a __syncthreads() inside orders the
  // barrier / TMEM set-up above as well)
  load_weights_tc_staged<WS_THREADS>(sw, prm, sw + 2 * FW_FLOATS);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *slot;

  const int S = P.S;
  const bool is_mlp = warp < 8;
  const int g = is_mlp ? (warp >> 2) : ((warp - 8) >> 3);           // sub-CTA
  const int wsub = is_mlp ? (warp & 3) : 4 + ((warp - 8) & 7);        // warp index inside the sub-CTA: 0..3 MLP, 4..11 gather
  const int tsub = wsub * 32 + lane;                                  // thread index inside the sub-CTA
  float* ring = sw + 2 * FW_FLOATS + g * (WS_NST * WS_STAGE_FLOATS);
  float* unit = sw + 2 * FW_FLOATS + 2 * (WS_NST * WS_STAGE_FLOATS) + g * (rpu * (6 + 6 * S));
  float* s_ray = unit;
  float* s_z = s_ray + rpu * 6;
  float* s_raw = s_z + rpu * S;
  const float2* grid = reinterpret_cast<const float2*>(prm.grid);
  double st_acc = lane == NRT_STAT_UNCERT_MIN ? 3.4e38 : 0.0;      // lane k: partial sum of loss statistic k (LossFuse)

  TileCtx c;
  c.tb = tmem_base + 256u * (uint32_t)g;
  c.lane_tb = c.tb + ((uint32_t)(32 * (warp & 3)) << 16);
  c.bar = &bars[g];
  c.phase = 0u;
  c.w_hi = smem_u32(sw);
  c.w_lo = smem_u32(sw + FW_FLOATS);
  const bool issuer = is_mlp && (warp & 3) == 0;

  if (tr) tr_pro = clock64() - tr_t0;
  const int64_t n_units = (n_rays + rpu - 1) / rpu;
  int cnt = 0;                                                        // tiles this sub-CTA has pushed through its ring
  int total_tiles = 0;                                                // ... and will have pushed at the end
  for (int64_t un = (int64_t)blockIdx.x * 2 + g; un < n_units; un += (int64_t)gridDim.x * 2) {
    const int nr = (int)min((int64_t)rpu, n_rays - un * rpu);
    total_tiles += (nr * S + WS_ROWS - 1) / WS_ROWS;
  }
  for (int64_t un = (int64_t)blockIdx.x * 2 + g; un < n_units; un += (int64_t)gridDim.x * 2) {
    const int64_t r0 = un * rpu;
    const int nr = (int)min((int64_t)rpu, n_rays - r0);
    const int npts = nr * S;
    const long long tr_a = tr ? clock64() : 0;
    // ---- stage the rays and their depth samples (all 12 warps) ----
    for (int i = tsub; i < nr * 6; i += WS_SUB) {
      const int rl = i / 6, k = i - rl * 6;
      s_ray[i] = k < 3 ? __ldg(rays_o + (r0 + rl) * 3 + k) : __ldg(rays_d + (r0 + rl) * 3 + k - 3);
    }
    for (int rl = wsub; rl < nr; rl += 12) {
      const int64_t ray = r0 + rl;
      float* z = s_z + rl * S;
      if (z_in) {
        for (int s = lane; s < S; s += 32) z[s] = __ldg(z_in + ray * S + s);
      } else {
        warp_sample_z(P, __ldg(target_d + ray), u ? u + ray * S : nullptr, perturb, seed, ray, z, lane);
      }
    }
    bar_sync(WS_BAR_UNIT(g), WS_SUB);
    const long long tr_b = tr ? clock64() : 0;
    // ---- tiles of 128 points ----
    if (is_mlp) {
      const int row = tsub;                                           // 0..127 = TMEM lane
      for (int t0 = 0; t0 < npts; t0 += WS_ROWS, ++cnt) {
        const int st = cnt % WS_NST;
        const int pl = t0 + row;
        float x0, x1, x2;
        ws_point(P, s_ray, s_z, pl, npts, S, x0, x1, x2);
        // OneBlob and the uncertainty sample do not depend on the gather: do them while the stage fills
#pragma unroll 1
        for (int d = 0; d < 3; ++d) {
          float bins[NRT_BINS];
          oneblob16_fast(d == 0 ? x0 : d == 1 ? x1 : x2, bins);
          stage16(c, TA_OB + 16 * d, bins);
        }
        const float unc = uncert_sample(P, prm.uncert, x0, x1, x2);
        bar_sync(WS_BAR_FULL(g, st), WS_SUB);
        {
          const float4* sf = reinterpret_cast<const float4*>(ring + st * WS_STAGE_FLOATS);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            float f[16];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float4 v = sf[(4 * q + k) * WS_ROWS + row];
              f[4 * k] = v.x;
              f[4 * k + 1] = v.y;
              f[4 * k + 2] = v.z;
              f[4 * k + 3] = v.w;
            }
            stage16(c, TA_X0 + 16 * q, f);
          }
        }
        if (cnt + WS_NST < total_tiles) bar_arrive(WS_BAR_EMPTY(g, st), WS_SUB);     // someone will refill this stage
        // ---- phase 1: h1 = relu(W1 [hash | oneblob]) ----
        ws_run_layer<80, 32>(c, g, issuer, TA_X0, c.w_hi + FW_W1 * 4, c.w_lo + FW_W1 * 4);
        uint32_t m1 = 0u, m3 = 0u;                                     // ReLU masks, saved for the backward pass
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          float h[16];
          tmem_ld16(c.lane_tb + TC_ACC + 16 * q, h);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) h[j] = fmaxf(h[j], 0.f);
          if (TRAIN && out.masks) {
#pragma unroll
            for (int j = 0; j < 16; ++j) m1 |= (h[j] > 0.f ? 1u : 0u) << (16 * q + j);
          }
          stage16(c, TA_X0 + 16 * q, h);
        }
        // ---- phase 2 on [h1 | oneblob]: o = W2 h1 (columns 0..15) and a3 = W23 h1 + W3_ob oneblob (columns 16..47) ----
        ws_run_layer<80, 48>(c, g, issuer, TA_X0, c.w_hi + FW_W23 * 4, c.w_lo + FW_W23 * 4);
        float o4[4];
        tmem_ld4(c.lane_tb + TC_ACC, o4);                              // o[0] = sdf
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          float h[16];
          tmem_ld16(c.lane_tb + TC_ACC + 16 + 16 * q, h);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) h[j] = fmaxf(h[j], 0.f);
          if (TRAIN && out.masks) {
#pragma unroll
            for (int j = 0; j < 16; ++j) m3 |= (h[j] > 0.f ? 1u : 0u) << (16 * q + j);
          }
          stage16(c, TA_X0 + 16 * q, h);
        }
        if (TRAIN && out.masks && pl < npts) reinterpret_cast<uint2*>(out.masks)[r0 * S + pl] = make_uint2(m1, m3);
        // ---- phase 3: rgb logits = W4 relu(a3) ----
        ws_run_layer<32, 16>(c, g, issuer, TA_X0, c.w_hi + FW_W4 * 4, c.w_lo + FW_W4 * 4);
        float r4[4];
        tmem_ld4(c.lane_tb + TC_ACC, r4);
        tmem_ld_wait();
        if (pl < npts) {
          float* r = s_raw + pl * 5;
          r[0] = r4[0];
          r[1] = r4[1];
          r[2] = r4[2];
          r[3] = o4[0];
          r[4] = unc;
        }
      }
    } else {
      const int j = wsub - 4;                                         // gather warp 0..7
      const int row = 32 * (j & 3) + lane;
      const int lbase = 8 * (j >> 2);
      for (int t0 = 0; t0 < npts; t0 += WS_ROWS, ++cnt) {
        const int st = cnt % WS_NST;
        const int pl = t0 + row;
        float x0, x1, x2;
        ws_point(P, s_ray, s_z, pl, npts, S, x0, x1, x2);
        const bool save_feat = TRAIN && out.feat && pl < npts;
        const int64_t pglob = r0 * S + pl;                               // flat point index of this row
        if (cnt >= WS_NST) bar_sync(WS_BAR_EMPTY(g, st), WS_SUB);
        float4* sf = reinterpret_cast<float4*>(ring + st * WS_STAGE_FLOATS);
#pragma unroll 1
        for (int b = 0; b < 4; ++b) {
          const int l0 = lbase + 2 * b;
          float f[4];
          gather_levels_paired<2>(P.lv + l0, grid, x0, x1, x2, f);
          const float4 v = make_float4(f[0], f[1], f[2], f[3]);
          sf[(l0 >> 1) * WS_ROWS + row] = v;
          if (save_feat) reinterpret_cast<float4*>(out.feat)[feat_tiled_index(pglob, l0 >> 1)] = v;
        }
        bar_arrive(WS_BAR_FULL(g, st), WS_SUB);
      }
    }
    bar_sync(WS_BAR_UNIT(g), WS_SUB);
    const long long tr_c = tr ? clock64() : 0;
    // ---- integrate along each ray (all 12 warps) ----
    for (int rl = wsub; rl < nr; rl += 12) {
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
      if (TRAIN && lf.target_rgb) {
        // per-sample masks: front = z < d - tr ; back = z > d + tr ; sdf_mask = !front & !back & (d > 0)   (tp/model/utils.py:81-148)
        const float td = __ldg(lf.target_d + ray);
        const float tr = P.sc_trunc;
        const float lo = __fsub_rn(td, tr), hi = __fadd_rn(td, tr);
        float nfs = 0.f, fsq = 0.f, nsd = 0.f, ssq = 0.f;
        for (int s = lane; s < S; s += 32) {
          const float zz = z[s], sdf = raw[s * 5 + 3];
          if (zz < lo) {
            const float e = sdf - 1.0f;
            nfs += 1.0f;
            fsq = fmaf(e, e, fsq);
          } else if (!(zz > hi) && td > 0.0f) {
            const float e = __fadd_rn(zz, __fmul_rn(sdf, tr)) - td;
            nsd += 1.0f;
            ssq = fmaf(e, e, ssq);
          }
        }
        nfs = warp_sum(nfs);
        fsq = warp_sum(fsq);
        nsd = warp_sum(nsd);
        ssq = warp_sum(ssq);
        {
          const float e0 = ro.rgb[0] - __ldg(lf.target_rgb + ray * 3 + 0), e1 = ro.rgb[1] - __ldg(lf.target_rgb + ray * 3 + 1),
                      e2 = ro.rgb[2] - __ldg(lf.target_rgb + ray * 3 + 2);
          const bool valid = td > 0.0f && td < P.depth_trunc;
          const float ed = ro.depth - td;
          // every lane holds the ray's values (warp-uniform); lane k adds the one that belongs to statistic k
          double v = 0.0;
          switch (lane) {
            case NRT_STAT_N_RAYS: v = 1.0; break;
            case NRT_STAT_N_VALID: v = valid ? 1.0 : 0.0; break;
            case NRT_STAT_N_FS: v = (double)nfs; break;
            case NRT_STAT_N_SDF: v = (double)nsd; break;
            case NRT_STAT_N_SAMPLES: v = (double)S; break;
            case NRT_STAT_RGB_SQ: v = (double)(e0 * e0) + (double)(e1 * e1) + (double)(e2 * e2); break;
            case NRT_STAT_DEPTH_SQ: v = valid ? (double)(ed * ed) : 0.0; break;
            case NRT_STAT_FS_SQ: v = (double)fsq; break;
            case NRT_STAT_SDF_SQ: v = (double)ssq; break;
            case NRT_STAT_INV2U: v = valid ? (double)(1.0f / (2.0f * (ro.uncert + 1e-9f))) : 0.0; break;
            case NRT_STAT_LOGU: v = valid ? (double)logf(ro.uncert + 1e-9f) : 0.0; break;
            default: break;
          }
          st_acc = lane == NRT_STAT_UNCERT_MIN ? fmin(st_acc, (double)ro.uncert) : st_acc + v;
        }
      }
    }
    bar_sync(WS_BAR_UNIT(g), WS_SUB);
    if (tr) {
      const long long tr_d = clock64();
      tr_stage += tr_b - tr_a, tr_tiles += tr_c - tr_b, tr_comp += tr_d - tr_c;
    }
  }
  if (tr && g == 0 && tsub == 0) {
    long long* o = g_ws_trace + blockIdx.x * 8;
    o[0] = tr_pro, o[1] = tr_stage, o[2] = tr_tiles, o[3] = tr_comp, o[4] = clock64() - tr_t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<WS_COLS>(tmem_base);
  if (TRAIN && lf.target_rgb) {
    // fold the 24 warps in warp order, publish this CTA's partial, let the last CTA reduce all partials in CTA order
    volatile uint32_t* s_last = reinterpret_cast<volatile uint32_t*>(smem_raw + 32);    // header slack
    double* all = reinterpret_cast<double*>(sw + 2 * FW_FLOATS);                        // the ring area is idle by now
    if (lane < WS_STAT_SLOTS) all[warp * WS_STAT_SLOTS + lane] = st_acc;
    __syncthreads();
    if (threadIdx.x < NRT_N_STATS) {
      double v = 0.0;
      if (threadIdx.x < NRT_N_STATS_SUM) {
        for (int w = 0; w < WS_THREADS / 32; ++w) v += all[w * WS_STAT_SLOTS + threadIdx.x];
      } else if (threadIdx.x == NRT_STAT_UNCERT_MIN) {
        v = 3.4e38;
        for (int w = 0; w < WS_THREADS / 32; ++w) v = fmin(v, all[w * WS_STAT_SLOTS + threadIdx.x]);
      }
      lf.part[(int64_t)blockIdx.x * NRT_N_STATS + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) *s_last = atomicAdd(lf.counter, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (*s_last) {
      // The last CTA folds the per-CTA partials: 16 groups of 16 threads take every 16th CTA each (all their loads in flight at
      // once), then one thread per statistic adds the 16 group sums in group order -- a fixed tree, so the result does not
      // depend on scheduling.  (A single thread per statistic walking all CTAs in a dependent chain cost ~13 us of exposed L2
      // latency at the end of every training forward.)
      __threadfence();
      double* grp = all;                          // [16 groups][16 stats], the ring area is idle
      if (threadIdx.x < 256) {
        const int k = threadIdx.x & 15, j = threadIdx.x >> 4;
        double v = k == NRT_STAT_UNCERT_MIN ? 3.4e38 : 0.0;
        for (unsigned b = j; b < gridDim.x; b += 16) {
          const double pv = lf.part[(int64_t)b * NRT_N_STATS + k];
          v = k == NRT_STAT_UNCERT_MIN ? fmin(v, pv) : v + pv;
        }
        grp[j * 16 + k] = v;
      }
      __syncthreads();
      if (threadIdx.x < NRT_N_STATS) {
        const int k = threadIdx.x;
        double v = k == NRT_STAT_UNCERT_MIN ? 3.4e38 : 0.0;
#pragma unroll
        for (int j = 0; j < 16; ++j) v = k == NRT_STAT_UNCERT_MIN ? fmin(v, grp[j * 16 + k]) : v + grp[j * 16 + k];
        lf.stats[k] = v;
        if (k == 0) *lf.counter = 0u;   // re-arm for the next launch
      }
    }
    if (tr && threadIdx.x == 0) g_ws_trace[blockIdx.x * 8 + 5] = clock64() - tr_t0;
    if (lf.losses) {                            // one shard = the whole batch: the losses follow at once (no extra launch)
      __syncthreads();
      if (*s_last && threadIdx.x == 0) {
        __threadfence();
        finalize_losses(lf.stats, lf.losses);
      }
    }
  }
}

int launch_render_fwd_ws(const NrtPlan* plan, const NrtParams* prm, const float* rays_o, const float* rays_d,
                         const float* target_d, int64_t n_rays, const float* z_in, const float* u, int perturb, uint64_t seed,
                         const NrtRenderOut* out, const float* target_rgb, double* stats, float* losses, const int* seed_step,
                         cudaStream_t st) {
  if (n_rays == 0) return NRT_OK;
  const int S = plan->dev.S;
  LossFuse lf{};
  static const int fwd_dbg = [] {
    const char* e = getenv("NRT_FWD_DEBUG");
    return e ? atoi(e) : 0;
  }();
  lf.trace = fwd_dbg;
  if (target_rgb) {          // stats buffer layout as launch_loss_partial (forward.cu): [stats(16) | counter | CTA partials]
    lf.target_rgb = target_rgb;
    lf.target_d = target_d;
    lf.stats = stats;
    lf.counter = reinterpret_cast<unsigned int*>(stats + NRT_N_STATS);
    lf.part = stats + NRT_N_STATS + 2;
    lf.losses = losses;
  }
  const int64_t slots = 2 * (int64_t)plan->sm_count;
  // rays per block: every sub-CTA gets the same number of blocks (rounds), each block as large as the staging buffer allows
  const int cap = WS_UNIT_PTS / S > 1 ? WS_UNIT_PTS / S : 1;
  const int64_t rounds = (n_rays + slots * cap - 1) / (slots * cap);
  int64_t rpu = (n_rays + slots * rounds - 1) / (slots * rounds);
  if (rpu > cap) rpu = cap;
  if (rpu < 1) rpu = 1;
  const int64_t units = (n_rays + rpu - 1) / rpu;
  const size_t smem = TC_SMEM_WEIGHTS + (size_t)(2 * WS_NST * WS_STAGE_FLOATS + 2 * rpu * (6 + 6 * S)) * sizeof(float);
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    NRT_CUDA_CHECK(cudaFuncSetAttribute(render_fwd_ws_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    NRT_CUDA_CHECK(cudaFuncSetAttribute(render_fwd_ws_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  const int64_t want = (units + 1) / 2;
  const int blocks = (int)(want < plan->sm_count ? want : plan->sm_count);
  if (out->feat || out->masks || target_rgb)
    render_fwd_ws_kernel<true><<<blocks, WS_THREADS, smem, st>>>(plan->dev, *prm, rays_o, rays_d, target_d, n_rays, z_in, u, perturb,
                                                                seed, seed_step, (int)rpu, *out, lf);
  else
    render_fwd_ws_kernel<false><<<blocks, WS_THREADS, smem, st>>>(plan->dev, *prm, rays_o, rays_d, target_d, n_rays, z_in, u, perturb,
                                                                 seed, seed_step, (int)rpu, *out, lf);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}

int ws_trace_read(void* dst, int bytes) {
  const int a = (int)sizeof(long long) * 256 * 8;
  NRT_CUDA_CHECK(cudaMemcpyFromSymbol(dst, g_ws_trace, bytes < a ? bytes : a));
  return NRT_OK;
}
