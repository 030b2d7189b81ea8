// Tensor-core backward of the two MLPs (what autograd does between dL/d raw and dL/d hash-features / dL/d weights for
// loss.backward() at src/slam/coslam/coslam.py:216,368 through src/slam/coslam/model/decoder.py:29-41,99-116).
//
// One persistent CTA per SM, 256 threads, thread pair (r, r+128) = point r of the tile = TMEM lane r, each thread owning
// half of the columns of every activation / gradient (mlp_tc.cuh).  Per tile of 128 points:
//   recompute   h1 = relu(W1 [hash|oneblob]) ; [o | a3] = [W2 h1 | W23 h1 + W3_ob oneblob], h3 = relu(a3)   (2 tcgen05 phases)
//   data grads  dh3 = W4^T dc (SIMT, 96 FMA) ; [dh1 | do] from [da3 | dsdf] ; dfeat = W1h^T da1               (2 tcgen05 phases)
//               (layer fusion through W23 = W3_geo W2_geo, mlp_tc.cuh: four dependent phases per tile instead of six; 3xTF32)
//   weight grads: every activation X = [hash|oneblob|geo|h1|h3] (160) and every upstream gradient Y = [da1|da3|do|dc] (88)
//               is also scattered, transposed and rounded to tf32, into shared memory as K-major operands over the point
//               index, and ONE tcgen05 GEMM  D[128 x 160] += Y^T X  (K = 128 points, 16 MMAs) accumulates all four weight
//               gradients as sub-blocks of a TMEM-resident accumulator that lives across all tiles of the CTA:
//                   dW1 = D[0:32, 0:80]   dW3 = D[32:64, 32:80 | 81:96]   dW2 = D[64:80, 96:128]   dW4 = D[80:83, 128:160]
//               (the other blocks are by-products the tensor pipe computes for free); flushed once per CTA with atomics.
//   table grads: the CTA carries eight more warps (threads 0..255; the MLP group is threads 256..511) that take each finished tile of dL/d hash-features from
//               a shared-memory hand-over slot and scatter it into the table gradient with red.global.add.v2.f32 while the
//               MLP warps work on the next tile, so the RED-bound scatter overlaps the latency-bound MLP chain and the
//               feature gradients never travel through HBM.  Scatter thread = (point, 8 levels); on coarse levels runs of
//               consecutive samples that fall into the same cell are merged with warp shuffles before the reduction.
// TMEM: [0,32) accumulator | [32,128) A_hi | [128,224) A_lo | [224,384) D.   Shared memory ~221 KB.
#include "mlp_tc.cuh"

#define BT_COLS 512
#define TC_DW 224

// backward weight block (floats), hi at +0 and lo at +BW_FLOATS
// phase B12 on A = [da3 (32) | dsdf, 0 x 7] (K = 40): columns 0..31 = dh1 = W23^T da3 + W2[0,:]^T dsdf,
//                                                   columns 32..47 = d o = [dsdf, W3[:, 48:63]^T da3]
#define BW_B12 0                   // [10 K-chunks][48 rows][4]
#define BW_W1T (BW_B12 + 40 * 48)  // dfeat = da1 W1[:, :32]:  [8 K-chunks j][32 rows f][4]
#define BW_FLOATS (BW_W1T + 32 * 32)

// transposed operands of the weight-gradient GEMM: buf[32 chunks of 4 points][R rows][4 points], R = 1 (mod 8)
#define XT_ROWS 161                // hash 0..31 | oneblob 32..79 | o = [sdf, geo] 80..95 | h1 96..127 | h3 128..159
#define YT_ROWS 89                 // da1 0..31 | da3 32..63 | do 64..79 | dc 80..82 | zero 83..88
#define XR_HASH 0
#define XR_OB 32
#define XR_GEO 80
#define XR_H1 96
#define XR_H3 128
#define YR_DA1 0
#define YR_DA3 32
#define YR_DO 64
#define YR_DC 80
#define BWD_THREADS 512            // 256 scatter threads (warps 0..7) + 256 MLP threads (warps 8..15)
#define RING_DF_FLOATS (8 * 128 * 4)          // dfeat of one tile, chunk-major [8 chunks][128 rows][4]
#define RING_STAGE_FLOATS (RING_DF_FLOATS + 4 * 128)   // + x0[128] x1[128] x2[128] active[128]
#define BAR_FULL(s) (2 + 2 * (s))   // named barriers: hand-over stage s filled / drained
#define BAR_EMPTY(s) (3 + 2 * (s))
#define XT_FLOATS (32 * XT_ROWS * 4)
#define YT_FLOATS (32 * YT_ROWS * 4 + 160)     // the MMA reads 128 rows per chunk: slack behind the last chunk

__device__ __forceinline__ void put_split_bw(float* blk, int o, float v) {
  const float h = tf32_hi(v);
  blk[o] = h;
  blk[o + BW_FLOATS] = v - h;
}

// ---------------------------------------------------------------------------------------------
// scatter warps (threads 0..255): warp w takes levels w and w + 8 (one coarse, one fine) of all 128 rows of the tile, 32
// consecutive rows at a time, so every warp sees the same mix of samples in front of and behind the surface.
// On levels flagged `agg` consecutive rows (= consecutive samples of a ray) mostly share the trilinear cell: each lane
// folds the contributions of the following lanes of its window (8 lanes / 3 shuffle steps on the coarsest levels, 4 lanes /
// 2 steps on the medium ones; DevLevel::agg) that sit in the same cell, and only the first lane of each run issues the
// reductions.
// ---------------------------------------------------------------------------------------------
template <int NSTG>
__device__ __forceinline__ void scatter_warps(const DevPlan& P, const DevLevel* __restrict__ s_lv, const float* __restrict__ ring,
                                              float2* __restrict__ dgrid, int64_t n_tiles) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int my_tiles = blockIdx.x < n_tiles ? (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
  for (int k = 0; k < my_tiles; ++k) {
    const int stg = k % NSTG;
    const float* stage = ring + stg * RING_STAGE_FLOATS;
    bar_sync(BAR_FULL(stg), BWD_THREADS);
    const float* xs = stage + RING_DF_FLOATS;
    // iteration = (block of 32 consecutive rows, one of this warp's two levels)
#pragma unroll 1
    for (int it = 0; it < (dgrid ? 8 : 0); ++it) {
      const int row = 32 * (it >> 1) + lane;
      const int lg = warp + 8 * (it & 1);
      const float x0 = xs[row], x1 = xs[128 + row], x2 = xs[256 + row];
      const bool active = xs[384 + row] != 0.f;
      const DevLevel& L = s_lv[lg];
      const float2 g = *reinterpret_cast<const float2*>(stage + ((lg >> 1) * 128 + row) * 4 + (lg & 1) * 2);
      const bool nz = active && (g.x != 0.f || g.y != 0.f);
      if (!__any_sync(0xffffffffu, nz)) continue;          // e.g. 32 samples behind the surface: nothing to add
      uint32_t idx[8];
      float w[8];
      const LevelPos p = level_corners(L, x0, x1, x2, idx, w);
      float2* base = dgrid + L.offset;
      if (!L.agg) {
        if (nz) {
#pragma unroll
          for (int c = 0; c < 8; ++c) red_add_f2(base + idx[c], w[c] * g.x, w[c] * g.y);
        }
      } else {
        // every lane takes part in the shuffles; lanes without a gradient contribute zeros.
        // same[s]: the lane 2^s places further in this window of 2^agg lanes sits in the same cell
        const int wmask = (1 << L.agg) - 1;
        bool same[3];
#pragma unroll
        for (int sdx = 0; sdx < 3; ++sdx) {
          const int d = 1 << sdx;
          const uint32_t o0 = __shfl_down_sync(0xffffffffu, p.g[0], d), o1 = __shfl_down_sync(0xffffffffu, p.g[1], d),
                         o2 = __shfl_down_sync(0xffffffffu, p.g[2], d);
          same[sdx] = ((lane & wmask) + d <= wmask) && o0 == p.g[0] && o1 == p.g[1] && o2 == p.g[2];
        }
        const uint32_t q0 = __shfl_up_sync(0xffffffffu, p.g[0], 1), q1 = __shfl_up_sync(0xffffffffu, p.g[1], 1),
                       q2 = __shfl_up_sync(0xffffffffu, p.g[2], 1);
        const bool head = (lane & wmask) == 0 || !(q0 == p.g[0] && q1 == p.g[1] && q2 == p.g[2]);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float v0 = nz ? w[c] * g.x : 0.f, v1 = nz ? w[c] * g.y : 0.f;
#pragma unroll
          for (int sdx = 0; sdx < 3; ++sdx) {
            if (sdx < (int)L.agg) {          // warp-uniform: 2 steps on the medium levels, 3 on the coarsest
              const float a0 = __shfl_down_sync(0xffffffffu, v0, 1 << sdx), a1 = __shfl_down_sync(0xffffffffu, v1, 1 << sdx);
              if (same[sdx]) {
                v0 += a0;
                v1 += a1;
              }
            }
          }
          if (head && (v0 != 0.f || v1 != 0.f)) red_add_f2(base + idx[c], v0, v1);
        }
      }
    }
    if (k + NSTG < my_tiles) bar_arrive(BAR_EMPTY(stg), BWD_THREADS);
  }
}

// SAVED: the forward pass saved the ReLU masks of both hidden layers (NrtRenderOut::masks).  The network is piecewise linear,
// so with the masks given every DATA gradient (dfeat, d uncert) is independent of the recomputed activation values; those
// only feed the tf32-rounded operands of the weight-gradient GEMM.  The two recompute phases then run single-pass TF32 on the
// hi pieces alone: a third of the MMAs, no lo staging, no lo copy of W1 / W23 in shared memory -- which makes room for a
// second hand-over stage, so the MLP group no longer waits for the scatter warps to drain the previous tile.
template <bool SAVED>
__global__ void __launch_bounds__(BWD_THREADS, 1) decode_bwd_tc_kernel(const __grid_constant__ DevPlan P, const NrtParams prm,
                                                                      const PointSource src, int64_t n_pts,
                                                                      const float* __restrict__ feat, int feat_tiled,
                                                                      const uint32_t* __restrict__ masks,
                                                                      const float* __restrict__ draw, float* __restrict__ dfeat,
                                                                      const NrtGrads grads) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int NSTG = SAVED ? 2 : 1;
  float* rest;
  TileCtx c;
  if (SAVED) {
    // cta_prologue() with the hi pieces of W1 / W23 only
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    uint32_t* slot = reinterpret_cast<uint32_t*>(smem_raw + 8);
    float* sw = reinterpret_cast<float*>(smem_raw + TC_SMEM_HEADER);
    if ((threadIdx.x >> 5) == 0) {
      if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
      }
      __syncwarp();
      tmem_alloc<BT_COLS>(slot);
    }
    load_weights_hi(sw, prm);
    c.bar = bar;
    c.phase = 0u;
    c.w_hi = smem_u32(sw);
    c.w_lo = 0u;
    rest = sw + FW_W4;
  } else {
    c = cta_prologue<BT_COLS>(smem_raw, prm, &rest);
  }
  float* bw = rest;                       // backward weights hi | lo
  float* w4s = bw + 2 * BW_FLOATS;        // w4 as stored [3][32] (+ pad)
  float* xt = w4s + 128;
  float* yt = xt + XT_FLOATS;
  float* ring = yt + YT_FLOATS;
  __shared__ DevLevel s_lv[NRT_L];
  const int t = threadIdx.x, half = tc_half() & 1, row = tc_row();
  if (t < NRT_L) s_lv[t] = P.lv[t];
  for (int i = t; i < 48 * 40; i += BWD_THREADS) {  // phase B12 operand, see BW_B12
    const int n = i / 40, k = i % 40;
    float v = 0.f;
    if (n < 32) v = k < 32 ? w23_at(prm, k, n) : k == 32 ? __ldg(prm.w2 + n) : 0.f;
    else if (k < 32) v = n > 32 ? __ldg(prm.w3 + k * 63 + NRT_OB + (n - 33)) : 0.f;
    else v = (k == 32 && n == 32) ? 1.0f : 0.f;
    put_split_bw(bw, BW_B12 + ((k >> 2) * 48 + n) * 4 + (k & 3), v);
  }
  for (int i = t; i < 32 * 32; i += BWD_THREADS) {  // W1T[f][j] = w1[j][f]
    const int f = i >> 5, j = i & 31;
    put_split_bw(bw, BW_W1T + ((j >> 2) * 32 + f) * 4 + (j & 3), __ldg(prm.w1 + j * 80 + f));
  }
  for (int i = t; i < 128; i += BWD_THREADS) w4s[i] = i < 96 ? __ldg(prm.w4 + i) : 0.f;
  for (int i = t; i < YT_FLOATS; i += BWD_THREADS) yt[i] = 0.f;   // rows 83.. and the slack stay zero for the whole kernel
  // second mbarrier: completion of the weight-gradient GEMM, which nobody needs before the NEXT tile overwrites X^T / Y^T
  uint64_t* bar_wg = reinterpret_cast<uint64_t*>(smem_raw + 16);
  uint32_t ph_wg = 0u;
  if (t == 0) {
    mbar_init(bar_wg, 1);
    fence_mbar_init();
  }
  cta_prologue_finish(smem_raw, c);
  const uint32_t bw_hi = smem_u32(bw), bw_lo = smem_u32(bw + BW_FLOATS);
  const uint32_t xt_s = smem_u32(xt), yt_s = smem_u32(yt);
  // this row's column of the transposed operands
  float* xcol = xt + (row >> 2) * (XT_ROWS * 4) + (row & 3);
  float* ycol = yt + (row >> 2) * (YT_ROWS * 4) + (row & 3);

  const int64_t n_tiles = (n_pts + 127) / 128;
  bool first_tile = true;
  // Warp roles: the SM's issue arbiter favours high warp ids, so the latency-critical MLP chain takes warps 8..15 and
  // the throughput-bound scatter takes warps 0..7.
  if (t < TC_THREADS) {
    scatter_warps<NSTG>(P, s_lv, ring, reinterpret_cast<float2*>(grads.grid), n_tiles);
  } else {
  int k = 0;                      // tiles done by this CTA
  for (int64_t tl = blockIdx.x; tl < n_tiles; tl += gridDim.x, ++k) {
    const int64_t pt = tl * 128 + row;
    const bool active = pt < n_pts;
    float x0 = 0.f, x1 = 0.f, x2 = 0.f;
    float dc[3] = {0.f, 0.f, 0.f}, dsdf = 0.f, du = 0.f;
    unsigned m1 = 0u, m3 = 0u;        // ReLU masks of this half's 16 units of either hidden layer
    if (active) {
      if (SAVED) {
        const uint2 mk = __ldg(reinterpret_cast<const uint2*>(masks) + pt);
        m1 = (mk.x >> (16 * half)) & 0xffffu;
        m3 = (mk.y >> (16 * half)) & 0xffffu;
      }
      fetch_point(P, src, pt, x0, x1, x2);
      const float* g = draw + pt * 5;
      dc[0] = __ldg(g);
      dc[1] = __ldg(g + 1);
      dc[2] = __ldg(g + 2);
      if (half == 0) dsdf = __ldg(g + 3);
      else du = __ldg(g + 4);
    }
    // ---- saved hash features + OneBlob -> A operand (TMEM) and X^T (smem) ----
    {
      float f[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int64_t fi = feat_tiled ? feat_tiled_index(pt, half * 4 + q) : pt * 8 + half * 4 + q;
        const float4 v = active ? __ldg(reinterpret_cast<const float4*>(feat) + fi) : make_float4(0, 0, 0, 0);
        f[4 * q] = v.x;
        f[4 * q + 1] = v.y;
        f[4 * q + 2] = v.z;
        f[4 * q + 3] = v.w;
      }
      // the previous tile's weight-gradient GEMM still reads X^T / Y^T: wait for it here, with this tile's loads in flight
      if (k > 0) {
        __syncwarp();
        mbar_wait(bar_wg, ph_wg);
        ph_wg ^= 1u;
        tc_fence_after();
      }
      float hi[16], lo[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        hi[i] = tf32_hi(f[i]);
        lo[i] = f[i] - hi[i];
        xcol[(XR_HASH + 16 * half + i) * 4] = hi[i];
      }
      tmem_st16(c.lane_tb + TC_AHI + TA_X0 + 16 * half, hi);
      if (!SAVED) tmem_st16(c.lane_tb + TC_ALO + TA_X0 + 16 * half, lo);
    }
#pragma unroll 1
    for (int d = 2 * half; d < 2 + half; ++d) {
      float bins[NRT_BINS];
      oneblob16_fast(d == 0 ? x0 : d == 1 ? x1 : x2, bins);
      float hi[16], lo[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float v = active ? bins[i] : 0.f;
        hi[i] = tf32_hi(v);
        lo[i] = v - hi[i];
        xcol[(XR_OB + 16 * d + i) * 4] = hi[i];
      }
      tmem_st16(c.lane_tb + TC_AHI + TA_OB + 16 * d, hi);
      if (!SAVED) tmem_st16(c.lane_tb + TC_ALO + TA_OB + 16 * d, lo);
    }
    // ---- recompute layer 1 ----
    if (SAVED) run_layer_1p<80, 32>(c, TA_X0, c.w_hi + FW_W1 * 4);
    else run_layer<80, 32>(c, TA_X0, c.w_hi + FW_W1 * 4, c.w_lo + FW_W1 * 4);
    {
      float h[16];
      tmem_ld16(c.lane_tb + TC_ACC + 16 * half, h);
      tmem_ld_wait();
      float hi[16], lo[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float v = fmaxf(h[i], 0.f);
        if (!SAVED && v > 0.f) m1 |= 1u << i;
        hi[i] = tf32_hi(v);
        lo[i] = v - hi[i];
        xcol[(XR_H1 + 16 * half + i) * 4] = hi[i];
      }
      tmem_st16(c.lane_tb + TC_AHI + TA_X0 + 16 * half, hi);
      if (!SAVED) tmem_st16(c.lane_tb + TC_ALO + TA_X0 + 16 * half, lo);
    }
    // ---- recompute phase 2 on [h1 | oneblob]: o = [sdf, geo] (columns 0..15) and a3 (columns 16..47) ----
    if (SAVED) run_layer_1p<80, 48>(c, TA_X0, c.w_hi + FW_W23 * 4);
    else run_layer<80, 48>(c, TA_X0, c.w_hi + FW_W23 * 4, c.w_lo + FW_W23 * 4);
    {
      float o[8], a3[16];
      tmem_ld8(c.lane_tb + TC_ACC + 8 * half, o);
      tmem_ld16(c.lane_tb + TC_ACC + 16 + 16 * half, a3);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 8; ++k) xcol[(XR_GEO + 8 * half + k) * 4] = tf32_hi(o[k]);
      // h3 = relu(a3); dh3 = W4^T dc; da3 = dh3 * relu'
      if (half == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) ycol[(YR_DC + k) * 4] = tf32_hi(dc[k]);
      }
      float hi[16], lo[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int j = 16 * half + i;
        const float hv = fmaxf(a3[i], 0.f);
        xcol[(XR_H3 + j) * 4] = tf32_hi(hv);
        const float d = dc[0] * w4s[j] + dc[1] * w4s[32 + j] + dc[2] * w4s[64 + j];
        const float v = (SAVED ? ((m3 >> i) & 1u) != 0u : hv > 0.f) ? d : 0.f;
        hi[i] = tf32_hi(v);
        lo[i] = v - hi[i];
        ycol[(YR_DA3 + j) * 4] = hi[i];
      }
      tmem_st16(c.lane_tb + TC_AHI + TA_X0 + 16 * half, hi);
      tmem_st16(c.lane_tb + TC_ALO + TA_X0 + 16 * half, lo);
      if (half == 0) {
        // the dsdf block [dsdf, 0 x 7] goes behind da3, into the first eight (now dead) OneBlob columns
        float dh[8], dl[8];
        dh[0] = tf32_hi(dsdf);
        dl[0] = dsdf - dh[0];
#pragma unroll
        for (int i = 1; i < 8; ++i) dh[i] = dl[i] = 0.f;
        tmem_st8(c.lane_tb + TC_AHI + TA_OB, dh);
        tmem_st8(c.lane_tb + TC_ALO + TA_OB, dl);
      }
    }
    // ---- phase B12 on [da3 | dsdf]: dh1 (columns 0..31) and d o (columns 32..47) ----
    run_layer<40, 48>(c, TA_X0, bw_hi + BW_B12 * 4, bw_lo + BW_B12 * 4);
    {
      float dh[16], dov[8];
      tmem_ld16(c.lane_tb + TC_ACC + 16 * half, dh);
      tmem_ld8(c.lane_tb + TC_ACC + 32 + 8 * half, dov);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 8; ++i) ycol[(YR_DO + 8 * half + i) * 4] = tf32_hi(dov[i]);
      float hi[16], lo[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float v = (m1 >> i) & 1u ? dh[i] : 0.f;
        hi[i] = tf32_hi(v);
        lo[i] = v - hi[i];
        ycol[(YR_DA1 + 16 * half + i) * 4] = hi[i];
      }
      tmem_st16(c.lane_tb + TC_AHI + TA_X0 + 16 * half, hi);
      tmem_st16(c.lane_tb + TC_ALO + TA_X0 + 16 * half, lo);
    }
    // ---- dfeat = da1 W1[:, :32]  and the weight-gradient GEMM D += Y^T X over this tile's 128 points ----
    fence_async_smem();            // X^T / Y^T were written with generic st.shared; the MMA reads them through the async proxy
    layer_publish();
    if (tc_issuer_warp()) {
      tc_fence_after();
      if (elect_one()) {
      issue_layer<32, 32>(c, TA_X0, bw_hi + BW_W1T * 4, bw_lo + BW_W1T * 4);
      mma_commit(c.bar);                       // dfeat: waited for right below
      constexpr uint32_t idesc = idesc_tf32(128, 160, 0, 0);
      const uint64_t yd = smem_desc(yt_s, YT_ROWS * 16, 128), xd = smem_desc(xt_s, XT_ROWS * 16, 128);
#pragma unroll
      for (int ks = 0; ks < 16; ++ks)
        mma_tf32_ss(c.tb + TC_DW, yd + (uint64_t)(2 * YT_ROWS * ks), xd + (uint64_t)(2 * XT_ROWS * ks), idesc, !(first_tile && ks == 0));
      mma_commit(bar_wg);                      // weight gradients: waited for at the top of the next tile / before the flush
      }
    }
    first_tile = false;
    layer_wait(c);
    {
      float df[16];
      tmem_ld16(c.lane_tb + TC_ACC + 16 * half, df);
      tmem_ld_wait();
      if (active && dfeat) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          reinterpret_cast<float4*>(dfeat + pt * NRT_ENC + 16 * half)[q] = make_float4(df[4 * q], df[4 * q + 1], df[4 * q + 2], df[4 * q + 3]);
      }
      // hand the tile to the scatter warps through the ring
      const int stg = k % NSTG;
      float* stage = ring + stg * RING_STAGE_FLOATS;
      if (k >= NSTG) bar_sync(BAR_EMPTY(stg), BWD_THREADS);
#pragma unroll
      for (int q = 0; q < 4; ++q) sts4(stage + ((4 * half + q) * 128 + row) * 4, df[4 * q], df[4 * q + 1], df[4 * q + 2], df[4 * q + 3]);
      if (half == 0) {
        float* xs = stage + RING_DF_FLOATS;
        xs[row] = x0;
        xs[128 + row] = x1;
        xs[256 + row] = x2;
        xs[384 + row] = active ? 1.0f : 0.0f;
      }
      bar_arrive(BAR_FULL(stg), BWD_THREADS);
    }
    // uncertainty grid: raw[...,4] is the trilinear sample itself
    if (half == 1 && active && grads.uncert && du != 0.f) {
      UncertPos up = uncert_pos(P, x0, x1, x2);
#pragma unroll
      for (int cnr = 0; cnr < 8; ++cnr) {
        float w;
        int off = uncert_corner(P, up, cnr, w);
        if (off >= 0) atomicAdd(grads.uncert + off, w * du);
      }
    }
  }
    if (!first_tile) {                          // the last tile's weight-gradient GEMM
      __syncwarp();
      mbar_wait(bar_wg, ph_wg);
      tc_fence_after();
    }
  }  // MLP threads
  // ---- flush the weight gradients: the thread pair of row r holds row r of D, half 0 columns 0..79, half 1 80..159 ----
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (t >= TC_THREADS && !first_tile) {
    for (int cb = 80 * half; cb < 80 * half + 80; cb += 16) {
      float v[16];
      tmem_ld16(c.lane_tb + TC_DW + cb, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int col = cb + i;
        if (row < 32) {                       // dW1[j][k], k = hash 0..31 | oneblob 32..79
          if (col < 80 && grads.w1) atomicAdd(grads.w1 + row * 80 + col, v[i]);
        } else if (row < 64) {                // dW3[j][kk], kk = oneblob 0..47 (cols 32..79) | geo 48..62 (cols 81..95)
          if (grads.w3) {
            if (col >= 32 && col < 80) atomicAdd(grads.w3 + (row - 32) * 63 + (col - 32), v[i]);
            else if (col >= 81 && col < 96) atomicAdd(grads.w3 + (row - 32) * 63 + (col - 33), v[i]);
          }
        } else if (row < 80) {                // dW2[i][j]
          if (col >= 96 && col < 128 && grads.w2) atomicAdd(grads.w2 + (row - 64) * 32 + (col - 96), v[i]);
        } else if (row < 83) {                // dW4[i][j]
          if (col >= 128 && grads.w4) atomicAdd(grads.w4 + (row - 80) * 32 + (col - 128), v[i]);
        }
      }
    }
  }
  cta_epilogue<BT_COLS>(c);
}

size_t decode_bwd_tc_smem(bool saved) {
  const size_t fw = saved ? (size_t)FW_W4 : (size_t)2 * FW_FLOATS;
  return TC_SMEM_HEADER + (fw + 2 * BW_FLOATS + 128 + XT_FLOATS + YT_FLOATS + (saved ? 2 : 1) * RING_STAGE_FLOATS) * sizeof(float);
}

// feat_tiled: feat is NrtRenderOut::feat (tile-major, common.cuh: feat_tiled_index); 0: plain [n,32] (nrt_decode_bwd's scratch)
int launch_decode_bwd(const NrtPlan* plan, const NrtParams* prm, const PointSource& src, int64_t n_pts, const float* feat,
                      int feat_tiled, const uint32_t* masks, const float* draw, float* dfeat, const NrtGrads* grads, cudaStream_t st) {
  if (n_pts == 0) return NRT_OK;
  const bool saved = masks != nullptr;
  const size_t smem = decode_bwd_tc_smem(saved);
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    NRT_CUDA_CHECK(cudaFuncSetAttribute(decode_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)decode_bwd_tc_smem(true)));
    NRT_CUDA_CHECK(cudaFuncSetAttribute(decode_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)decode_bwd_tc_smem(false)));
  }
  const int64_t tiles = (n_pts + 127) / 128;
  const int blocks = (int)(tiles < plan->sm_count ? tiles : plan->sm_count);
  if (saved)
    decode_bwd_tc_kernel<true><<<blocks, BWD_THREADS, smem, st>>>(plan->dev, *prm, src, n_pts, feat, feat_tiled, masks, draw, dfeat, *grads);
  else
    decode_bwd_tc_kernel<false><<<blocks, BWD_THREADS, smem, st>>>(plan->dev, *prm, src, n_pts, feat, feat_tiled, nullptr, draw, dfeat, *grads);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}
