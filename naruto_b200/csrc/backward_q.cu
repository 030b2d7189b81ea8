// Backward of the two MLPs + hash-table scatter for the training path (what autograd does between dL/d raw and the parameter
// gradients for loss.backward() at src/slam/coslam/coslam.py:216,368 through src/slam/coslam/model/decoder.py:29-41,99-116 and
// the tcnn grid backward), second generation: the launch behind nrt_render_bwd when the forward saved features and ReLU masks.
//
// Round 1's kernel (backward_tc.cu, kept for callers without saved masks) ran ONE tile of 128 points per SM through a chain of
// four dependent tensor phases on 8 warps, with 8 more warps scattering: 25 % warp occupancy, every pipe idle (latency-bound).
// This kernel keeps the mathematics and changes the structure:
//
//   * 32 warps per SM.  Warps 16..31 = MLP group, FOUR threads per point (thread t, t+128, t+256, t+384 share TMEM lane t and
//     own a quarter of the columns of every activation / gradient each), so the per-thread instruction chain between two tensor
//     phases is half as long and twice as many warps hide it.  Warps 0..15 = scatter group, one hash level per warp.
//   * TMA.  The tile's inputs -- 16 KB of saved hash features (already stored tile-major = the chunk-major shared-memory image),
//     2.5 KB of dL/d raw, 1 KB of ReLU masks -- arrive by cp.async.bulk into a 3-slot shared-memory ring, one tile ahead,
//     completion on an mbarrier (expect_tx): no global-load latency inside the chain.  The same slot is recycled as the
//     hand-over buffer: the MLP threads overwrite the features they consumed with the tile's dL/d features, the scatter warps
//     drain it (mbarrier `ready`), then the producer refills it (mbarrier `free`).
//   * Smaller weight-gradient operands.  geo = W2[1:] h1 is linear in h1, so dW3[:, 48:63] = (da3^T h1) W2[1:]^T and
//     dW2[1:] = W3[:, 48:63]^T (da3^T h1): the GEMM D[128 x 144] += Y^T X over the tile's points takes X = [hash | oneblob |
//     h1 | h3] (144) and Y = [da1 | da3 | dsdf | dc] (68), no o / geo recompute, no `do` rows; the two small contractions with
//     W2 / W3 happen once per CTA at the end.  Phase 2 shrinks to N = 32 (a3 only), phase B12 to N = 32 (dh1 only).
//
// Per tile:   h1 = relu(W1 [hash|oneblob])                                   (1-pass TF32: feeds tf32 operands only)
//             a3 = W23 h1 + W3_ob oneblob, h3 = relu(a3)                     (1-pass)
//             dh3 = W4^T dc (SIMT), da3 = mask3 * dh3
//             dh1 = W23^T da3 + W2[0]^T dsdf, da1 = mask1 * dh1              (3xTF32: data-gradient path)
//             dfeat = W1[:, :32]^T da1                                       (3xTF32)  -> slot -> scatter warps -> red.global
//             D += Y^T X                                                     (1-pass, own mbarrier, overlaps the next tile)
// TMEM: [0,32) accumulator | [32,112) A_hi | [112,152) A_lo | [160,304) D | [320,352) dh3 | [352,368) dc hi / lo.   Shared memory ~213 KB.
#include <cstdlib>

#include "mlp_tc.cuh"

#define Q_THREADS 1024
#define Q_SCAT 512                  // threads 0..511: scatter warps (the SM's issue arbiter favours high warp ids -> MLP chain on top)
#define Q_MLP 512
#define Q_NSLOT 3
#define Q_COLS 512
#define Q_ACC 0
#define Q_AHI 32
#define Q_ALO 112
#define Q_DW 160
#define Q_ACC2 320                  // [320,352) dh3 (accumulators sit on 32-column boundaries)
#define Q_DCH 352                   // [352,360) dc (hi), [360,368) dc (lo): A operand of dh3 = W4^T dc
#define Q_DCL 360
#define Q_BAR_MLP 1                 // named barrier of the 512 MLP threads

// weights in shared memory (floats), chunk-major K-major B operands [K/4][rows][4]
#define QW_W1 0                          // [20][32][4]  k: hash 0..31 | oneblob 32..79          (tf32-rounded)
#define QW_A3 (QW_W1 + 80 * 32)          // [20][32][4]  k: h1 0..31 (W23) | oneblob 32..79 (W3)   (tf32-rounded)
#define QW_B12H (QW_A3 + 80 * 32)        // [10][32][4]  k: da3 0..31 (W23^T) | dsdf 32 (W2[0]) | 0
#define QW_B12L (QW_B12H + 40 * 32)
#define QW_W1TH (QW_B12L + 40 * 32)      // [8][32][4]   dfeat[f] = sum_j da1[j] w1[j][f]
#define QW_W1TL (QW_W1TH + 32 * 32)
#define QW_W4TH (QW_W1TL + 32 * 32)      // [2][32][4]   dh3[j] = sum_i dc[i] w4[i][j]  (k = i < 3, padded to 8)
#define QW_W4TL (QW_W4TH + 8 * 32)
#define QW_FLOATS (QW_W4TL + 8 * 32)

// transposed operands of the weight-gradient GEMM: buf[32 chunks of 4 points][R rows][4 points], R = 1 (mod 8)
#define QX_ROWS 145                      // hash 0..31 | oneblob 32..79 | h1 80..111 | h3 112..143
#define QY_ROWS 73                       // da1 0..31 | da3 32..63 | dsdf 64 | dc 65..67 | (garbage 68..)
#define QX_OB 32
#define QX_H1 80
#define QX_H3 112
#define QY_DA3 32
#define QY_DSDF 64
#define QY_DC 65
#define QY_LIVE 68                       // rows of D that carry weight gradients
#define QX_FLOATS (32 * QX_ROWS * 4)
#define QY_FLOATS (32 * QY_ROWS * 4 + (128 - QY_ROWS) * 4)   // the MMA reads 128 rows per chunk: slack behind the last chunk

// ring slot (floats): features in (chunk-major [8 chunks][128 rows][4], as saved by the forward) / feature gradients out
// (level-major [16 levels][128 rows][2])
#define QS_X 4096                        // x0[128] x1[128] x2[128]   (written by the MLP threads, read by the scatter warps)
#define QS_DRAW 4480                     // dL/d raw [128][5]         (TMA)
#define QS_MASK 5120                     // ReLU masks [128][2] words (TMA)
#define QS_FLOATS 5376
#define QS_TX_BYTES (16384 + 2560 + 1024)

#define Q_SMEM_HEADER 256
#define Q_SMEM_BYTES (Q_SMEM_HEADER + (QW_FLOATS + QX_FLOATS + QY_FLOATS + Q_NSLOT * QS_FLOATS) * 4)

// profiling aid (NRT_BWD_DEBUG bit 3): per-CTA clock64 stamps at the phase boundaries, read back with nrt_debug_read
__device__ long long g_q_trace[256 * 8];
__device__ long long g_q_warp[8 * 32 * 8];        // CTAs 0..7: per warp [loop, wait full/ready, wait mma, wait wg, publish, wait free]

namespace {

struct QBars {
  uint64_t mma;              // tensor phase complete
  uint64_t wg;               // weight-gradient GEMM complete
  uint64_t full[Q_NSLOT];    // TMA landed          (1 arrival + tx bytes)
  uint64_t free_[Q_NSLOT];   // scatter warps done  (16 arrivals)
  uint64_t ready[Q_NSLOT];   // dfeat written       (16 arrivals)
};

__device__ __forceinline__ void q_put_split(float* w, int hi_off, int lo_off, float v) {
  const float h = tf32_hi(v);
  w[hi_off] = h;
  w[lo_off] = v - h;
}

template <int K, int N>
__device__ __forceinline__ void q_issue_1p(uint32_t d, uint32_t a_hi, uint32_t wh) {
  constexpr uint32_t idesc = idesc_tf32(128, N, 0, 0);
  const uint64_t bh = smem_desc(wh, N * 16, 128);
#pragma unroll
  for (int ks = 0; ks < K / 8; ++ks) mma_tf32_ts(d, a_hi + 8 * ks, bh + (uint64_t)(2 * N * ks), idesc, ks > 0);
}
template <int K, int N>
__device__ __forceinline__ void q_issue_3p(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t wh, uint32_t wl) {
  constexpr uint32_t idesc = idesc_tf32(128, N, 0, 0);
  const uint64_t bh = smem_desc(wh, N * 16, 128), bl = smem_desc(wl, N * 16, 128);
#pragma unroll
  for (int ks = 0; ks < K / 8; ++ks) mma_tf32_ts(d, a_hi + 8 * ks, bh + (uint64_t)(2 * N * ks), idesc, ks > 0);
#pragma unroll
  for (int ks = 0; ks < K / 8; ++ks) mma_tf32_ts(d, a_lo + 8 * ks, bh + (uint64_t)(2 * N * ks), idesc, true);
#pragma unroll
  for (int ks = 0; ks < K / 8; ++ks) mma_tf32_ts(d, a_hi + 8 * ks, bl + (uint64_t)(2 * N * ks), idesc, true);
}

__device__ __forceinline__ void q_publish() {
  tmem_st_wait();
  tc_fence_before();
  bar_sync(Q_BAR_MLP, Q_MLP);
}
__device__ __forceinline__ void q_wait(uint64_t* bar, uint32_t& phase) {
  __syncwarp();
  mbar_wait(bar, phase);
  phase ^= 1u;
  tc_fence_after();
}

// TMA producer: one thread fills ring slot `slot` with tile `tl` (features, dL/d raw, masks)
__device__ __forceinline__ void q_prefetch(float* slot, uint64_t* full, int64_t tl, const float* __restrict__ feat,
                                           const float* __restrict__ draw, const uint32_t* __restrict__ masks) {
  fence_async_smem();                      // earlier generic-proxy accesses to the slot are ordered before the async-proxy writes
  mbar_arrive_expect_tx(full, QS_TX_BYTES);
  bulk_g2s(slot, feat + tl * 4096, 16384u, full);
  bulk_g2s(slot + QS_DRAW, draw + tl * 640, 2560u, full);
  bulk_g2s(slot + QS_MASK, masks + tl * 256, 1024u, full);
}

// ---------------------------------------------------------------------------------------------
// scatter warps (threads 0..511): warp w adds the share of level w (rows 0..63) and of level 15 - w (rows 64..127) of every
// finished tile into the table gradient, so every warp carries the same mix of coarse and fine levels.  On levels
// flagged `agg` consecutive rows (= consecutive samples of a ray) mostly share the trilinear cell: each lane folds the
// contributions of the following lanes of its window that sit in the same cell, and only the first lane of each run issues
// the reductions (same procedure as backward_tc.cu).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void q_scatter(const DevPlan& P, const float* __restrict__ slots, QBars* bars,
                                          int* __restrict__ s_cnt, float2* __restrict__ dgrid, int my_tiles, int64_t n_pts, int dbg) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int s = 0;
  uint32_t par = 0u;
  const bool wtrace = (dbg & 8) && blockIdx.x < 8;
  long long tw_ready = 0;
  const long long t_begin = wtrace ? clock64() : 0;
  for (int k = 0; k < my_tiles; ++k) {
    const float* slot = slots + s * QS_FLOATS;
    __syncwarp();
    const long long tw0 = wtrace ? clock64() : 0;
    mbar_wait(&bars->ready[s], par);
    if (wtrace) tw_ready += clock64() - tw0;
    const int64_t pt0 = ((int64_t)blockIdx.x + (int64_t)k * gridDim.x) * 128;
    // Work items of a tile = (level, block of 32 rows): 64 of them, handed out through a shared-memory counter, coarse
    // (run-merging, most expensive) levels first.  Which rows carry gradient depends on where the surface cuts the ray, so
    // any static split leaves some warps with twice the work of others (measured: 96 % vs 55 % busy).
    while (dgrid) {
      int item = 0;
      if (lane == 0) item = atomicAdd(s_cnt + s, 1);
      item = __shfl_sync(0xffffffffu, item, 0);
      if (item >= 4 * NRT_L) break;
      const int lg = item >> 2;
      const int row = 32 * (item & 3) + lane;
      // level constants straight from the kernel-parameter constant bank (warp-uniform index: constant-cache reads, no LSU
      // traffic -- the shared-memory copy cost 185 wavefronts per tile on the pipe that bounds this kernel)
      const DevLevel& L = P.lv[lg];
      const bool active = pt0 + row < n_pts;
      const float x0 = slot[QS_X + row], x1 = slot[QS_X + 128 + row], x2 = slot[QS_X + 256 + row];
      const float2 g = *reinterpret_cast<const float2*>(slot + (lg * 128 + row) * 2);      // level-major: 256 contiguous bytes per warp
      const bool nz = active && (g.x != 0.f || g.y != 0.f);
      if (!__any_sync(0xffffffffu, nz)) continue;          // e.g. 32 samples behind the surface: nothing to add
      uint32_t idx[8];
      float w[8];
      const LevelPos p = level_corners(L, x0, x1, x2, idx, w);
      float2* base = dgrid + L.offset;
      if (!L.agg || (dbg & 2)) {
        if (nz && !(dbg & 1)) {
#pragma unroll
          for (int c = 0; c < 8; c += 2) {
            if (dbg & 16) {
              red_add_f2(base + idx[c], w[c] * g.x, w[c] * g.y);
              red_add_f2(base + idx[c + 1], w[c + 1] * g.x, w[c + 1] * g.y);
            } else {
              red_add_xpair(base, idx[c], idx[c + 1], w[c] * g.x, w[c] * g.y, w[c + 1] * g.x, w[c + 1] * g.y);
            }
          }
        }
      } else {
        // every lane takes part in the shuffles; lanes without a gradient contribute zeros.
        // same[s]: the lane 2^s places further in this window of 2^agg lanes sits in the same cell
        const int wmask = (1 << L.agg) - 1;
        // one key per cell: the low 10 bits of each coordinate.  Lanes of one window hold consecutive samples of a ray, whose
        // cells are at most a few steps apart, so equal low bits <=> equal cell (coordinates differ by far less than 1024).
        const uint32_t key = (p.g[0] & 1023u) | ((p.g[1] & 1023u) << 10) | ((p.g[2] & 1023u) << 20);
        bool same[3];
#pragma unroll
        for (int sdx = 0; sdx < 3; ++sdx) {
          const int d = 1 << sdx;
          const uint32_t other = __shfl_down_sync(0xffffffffu, key, d);      // every lane shuffles (no short-circuit around it)
          same[sdx] = ((lane & wmask) + d <= wmask) && other == key;
        }
        const uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
        const bool head = (lane & wmask) == 0 || prev != key;
#pragma unroll
        for (int c = 0; c < 8; c += 2) {
          float v[4] = {nz ? w[c] * g.x : 0.f, nz ? w[c] * g.y : 0.f, nz ? w[c + 1] * g.x : 0.f, nz ? w[c + 1] * g.y : 0.f};
#pragma unroll
          for (int sdx = 0; sdx < 3; ++sdx) {
            if (sdx < (int)L.agg) {          // warp-uniform: 2 steps on the medium levels, 3 on the coarsest
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float a = __shfl_down_sync(0xffffffffu, v[e], 1 << sdx);
                if (same[sdx]) v[e] += a;
              }
            }
          }
          if (head && (v[0] != 0.f || v[1] != 0.f || v[2] != 0.f || v[3] != 0.f) && !(dbg & 1)) {
            if (dbg & 16) {
              red_add_f2(base + idx[c], v[0], v[1]);
              red_add_f2(base + idx[c + 1], v[2], v[3]);
            } else {
              red_add_xpair(base, idx[c], idx[c + 1], v[0], v[1], v[2], v[3]);
            }
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&bars->free_[s]);
    if (++s == Q_NSLOT) {
      s = 0;
      par ^= 1u;
    }
  }
  if (wtrace && lane == 0) {
    long long* o = g_q_warp + (blockIdx.x * 32 + warp) * 8;
    o[0] = clock64() - t_begin;
    o[1] = tw_ready;
  }
}

}  // namespace

__global__ void __launch_bounds__(Q_THREADS, 1) decode_bwd_q_kernel(const __grid_constant__ DevPlan P, const NrtParams prm,
                                                                    const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                                    const float* __restrict__ zv, int S, int64_t n_pts,
                                                                    const float* __restrict__ feat,
                                                                    const uint32_t* __restrict__ masks,
                                                                    const float* __restrict__ draw, const NrtGrads grads, float* __restrict__ wg_part,
                                                                    unsigned int* __restrict__ wg_counter, int dbg) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  QBars* bars = reinterpret_cast<QBars*>(smem_raw);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(smem_raw + 192);
  float* sw = reinterpret_cast<float*>(smem_raw + Q_SMEM_HEADER);
  float* xt = sw + QW_FLOATS;
  float* yt = xt + QX_FLOATS;
  float* slots = yt + QY_FLOATS;
  __shared__ int s_cnt[Q_NSLOT];                       // next scatter work item of the tile in each ring slot
  const int t = threadIdx.x;
  const int64_t n_tiles = (n_pts + 127) / 128;
  const int my_tiles = blockIdx.x < n_tiles ? (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

  const bool trace = (dbg & 8) && blockIdx.x < 256;
  if (trace && t == 0) g_q_trace[blockIdx.x * 8 + 0] = clock64();
  // ---- prologue: barriers, TMEM, the first tile's TMA (it flies while the weights are staged), weights ----
  if ((t >> 5) == 0) {
    if (t == 0) {
      mbar_init(&bars->mma, 1);
      mbar_init(&bars->wg, 1);
#pragma unroll
      for (int s = 0; s < Q_NSLOT; ++s) {
        mbar_init(&bars->full[s], 1);
        mbar_init(&bars->free_[s], Q_SCAT / 32);
        mbar_init(&bars->ready[s], Q_MLP / 32);
      }
      fence_mbar_init();
      if (my_tiles > 0) q_prefetch(slots, &bars->full[0], blockIdx.x, feat, draw, masks);
    }
    __syncwarp();
    tmem_alloc<Q_COLS>(tslot);
  }
  // raw weights -> shared memory (coalesced, one round trip), W23 = W3[:, 48:63] W2[1:16, :] formed ONCE per CTA from there
  // (the X^T region is free until the first tile): every operand image below is then built from shared memory
  float* raw1 = xt;                  // w1 [32][80]
  float* raw2 = raw1 + 2560;         // w2 [16][32]
  float* raw3 = raw2 + 512;          // w3 [32][63]
  float* w23s = raw3 + 2016 + 32;    // W23 [32][32]
  {
    // all seven loads of a thread in flight before the first store (one trip to L2 / HBM instead of seven)
    const float a0 = __ldg(prm.w1 + t), a1 = __ldg(prm.w1 + 1024 + t), a2 = t < 512 ? __ldg(prm.w1 + 2048 + t) : 0.f;
    const float b0 = t < 512 ? __ldg(prm.w2 + t) : 0.f;
    const float c0 = __ldg(prm.w3 + t), c1 = t < 2016 - 1024 ? __ldg(prm.w3 + 1024 + t) : 0.f;
    const float d0 = t < 96 ? __ldg(prm.w4 + t) : 0.f;
    raw1[t] = a0;
    raw1[1024 + t] = a1;
    if (t < 512) {
      raw1[2048 + t] = a2;
      raw2[t] = b0;
    }
    raw3[t] = c0;
    if (t < 2016 - 1024) raw3[1024 + t] = c1;
    if (t < 96) raw3[2016 + 32 + 1024 + t] = d0;          // w4 [3][32] behind W23
  }
  __syncthreads();
  if (trace && t == 0) g_q_trace[blockIdx.x * 8 + 6] = clock64();
  {
    const int j = t >> 5, m = t & 31;                        // W23[j][m] = sum_g w3[j][48+g] * w2[1+g][m]   (as w23_at)
    float acc = 0.f;
#pragma unroll
    for (int g = 0; g < NRT_GEO; ++g) acc = fmaf(raw3[j * 63 + NRT_OB + g], raw2[(1 + g) * 32 + m], acc);
    w23s[t] = acc;
  }
  __syncthreads();
  if (trace && t == 0) g_q_trace[blockIdx.x * 8 + 7] = clock64();
  for (int i = t; i < 80 * 32; i += Q_THREADS) {            // W1[j][k]
    const int j = i / 80, k = i % 80;
    sw[QW_W1 + ((k >> 2) * 32 + j) * 4 + (k & 3)] = tf32_hi(raw1[i]);
  }
  for (int i = t; i < 32 * 80; i += Q_THREADS) {            // a3 rows: [W23 | W3_ob]
    const int n = i / 80, k = i % 80;
    const float v = k < 32 ? w23s[n * 32 + k] : raw3[n * 63 + (k - 32)];
    sw[QW_A3 + ((k >> 2) * 32 + n) * 4 + (k & 3)] = tf32_hi(v);
  }
  for (int i = t; i < 32 * 40; i += Q_THREADS) {            // dh1[n] = sum_k da3[k] W23[k][n] + dsdf W2[0][n]
    const int n = i / 40, k = i % 40;
    const float v = k < 32 ? w23s[k * 32 + n] : k == 32 ? raw2[n] : 0.f;
    const int o = ((k >> 2) * 32 + n) * 4 + (k & 3);
    q_put_split(sw, QW_B12H + o, QW_B12L + o, v);
  }
  for (int i = t; i < 32 * 32; i += Q_THREADS) {            // W1T[f][j] = w1[j][f]
    const int f = i >> 5, j = i & 31;
    const int o = ((j >> 2) * 32 + f) * 4 + (j & 3);
    q_put_split(sw, QW_W1TH + o, QW_W1TL + o, raw1[j * 80 + f]);
  }
  for (int i = t; i < 8 * 32; i += Q_THREADS) {            // W4T[j][k] = w4[k][j] (k < 3)
    const int j = i >> 3, k = i & 7;
    const int o = ((k >> 2) * 32 + j) * 4 + (k & 3);
    q_put_split(sw, QW_W4TH + o, QW_W4TL + o, k < 3 ? w23s[1024 + k * 32 + j] : 0.f);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *tslot;
  bool any_tile = false;
  if (trace && t == 0) g_q_trace[blockIdx.x * 8 + 1] = clock64();

  if (t < Q_SCAT) {
    q_scatter(P, slots, bars, s_cnt, reinterpret_cast<float2*>(grads.grid), my_tiles, n_pts, dbg);
    if (trace && t == 0) g_q_trace[blockIdx.x * 8 + 3] = clock64();
  } else {
    // ------------------------------------------------------------------------------------------
    // MLP group
    // ------------------------------------------------------------------------------------------
    const int tm = t - Q_SCAT;
    const int row = tm & 127;
    const int q = __shfl_sync(0xffffffffu, tm >> 7, 0);          // quarter of the columns this thread owns (warp-uniform)
    const bool issuer = (tm >> 5) == 0;
    const uint32_t lane_tb = tb + ((uint32_t)(32 * ((tm >> 5) & 3)) << 16);
    const uint32_t w_s = smem_u32(sw), xt_s = smem_u32(xt), yt_s = smem_u32(yt);
    float* xcol = xt + (row >> 2) * (QX_ROWS * 4) + (row & 3);
    float* ycol = yt + (row >> 2) * (QY_ROWS * 4) + (row & 3);
    uint32_t ph_mma = 0u, ph_wg = 0u, par = 0u;
    int s = 0;
    const bool wtrace = (dbg & 8) && blockIdx.x < 8;
    long long tw_full = 0, tw_mma = 0, tw_wg = 0, tw_pub = 0, tw_free = 0;
    const long long t_begin = wtrace ? clock64() : 0;
#define QT0 const long long tq0 = wtrace ? clock64() : 0
#define QT1(acc) if (wtrace) acc += clock64() - tq0
    // ray of the tile's first point, advanced incrementally (no 64-bit division per tile)
    const uint32_t uS = (uint32_t)S;
    const uint32_t step = gridDim.x * 128u;
    const uint32_t step_ray = step / uS, step_rem = step % uS;
    int64_t ray0 = (int64_t)((blockIdx.x * 128u) / uS);
    uint32_t rem0 = (blockIdx.x * 128u) % uS;
    const float invS = 1.0f / (float)S;

    for (int k = 0; k < my_tiles; ++k) {
      const int64_t tl = (int64_t)blockIdx.x + (int64_t)k * gridDim.x;
      float* slot = slots + s * QS_FLOATS;
      const int64_t pt = tl * 128 + row;
      const bool active = pt < n_pts;
      // ---- this thread's coordinate(s): q < 3 needs x_q (OneBlob of dim q), q == 3 all three (uncertainty-grid gradient) ----
      float xq = 0.f, y0 = 0.f, y1 = 0.f, y2 = 0.f;
      if (active) {
        // (rem0 + row) / S for values < 384: exact through the float reciprocal (distance to the next integer >= 0.5 / S)
        const uint32_t v = rem0 + (uint32_t)row;
        const int64_t ray = ray0 + (int64_t)__float2uint_rz(((float)v + 0.5f) * invS);
        const float zz = __ldg(zv + pt);
        if (q < 3) {
          xq = normalise1(P, q, __fadd_rn(__ldg(rays_o + ray * 3 + q), __fmul_rn(__ldg(rays_d + ray * 3 + q), zz)));
        } else {
          y0 = normalise1(P, 0, __fadd_rn(__ldg(rays_o + ray * 3 + 0), __fmul_rn(__ldg(rays_d + ray * 3 + 0), zz)));
          y1 = normalise1(P, 1, __fadd_rn(__ldg(rays_o + ray * 3 + 1), __fmul_rn(__ldg(rays_d + ray * 3 + 1), zz)));
          y2 = normalise1(P, 2, __fadd_rn(__ldg(rays_o + ray * 3 + 2), __fmul_rn(__ldg(rays_d + ray * 3 + 2), zz)));
        }
      }
      // ---- the tile's inputs (TMA) ----
      __syncwarp();
      {
        QT0;
        mbar_wait(&bars->full[s], par);
        QT1(tw_full);
      }
      unsigned m1 = 0u, m3 = 0u;
      float dc[3] = {0.f, 0.f, 0.f}, dsdf = 0.f, du = 0.f;
      float f[8];
      {
        const float4 fa = *reinterpret_cast<const float4*>(slot + ((2 * q) * 128 + row) * 4);
        const float4 fb = *reinterpret_cast<const float4*>(slot + ((2 * q + 1) * 128 + row) * 4);
        f[0] = fa.x, f[1] = fa.y, f[2] = fa.z, f[3] = fa.w, f[4] = fb.x, f[5] = fb.y, f[6] = fb.z, f[7] = fb.w;
      }
      if (active) {
        const uint2 mk = reinterpret_cast<const uint2*>(slot + QS_MASK)[row];
        m1 = (mk.x >> (8 * q)) & 0xffu;
        m3 = (mk.y >> (8 * q)) & 0xffu;
        const float* g = slot + QS_DRAW + row * 5;
        dc[0] = g[0], dc[1] = g[1], dc[2] = g[2];
        dsdf = g[3];
        du = g[4];
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = 0.f;
      }
      if (q < 3) slot[QS_X + q * 128 + row] = xq;              // for the scatter warps
      // the previous tile's weight-gradient GEMM still reads X^T / Y^T
      if (k > 0) {
        QT0;
        q_wait(&bars->wg, ph_wg);
        QT1(tw_wg);
      }
      // ---- hash features -> A (TMEM) and X^T ----
      {
        float hi[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          hi[i] = tf32_hi(f[i]);
          xcol[(8 * q + i) * 4] = hi[i];
        }
        tmem_st8(lane_tb + Q_AHI + 8 * q, hi);
      }
      if (q < 3) {
        float bins[NRT_BINS];
        oneblob16_fast(xq, bins);
        float hi[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          hi[i] = tf32_hi(active ? bins[i] : 0.f);
          xcol[(QX_OB + 16 * q + i) * 4] = hi[i];
        }
        tmem_st16(lane_tb + Q_AHI + 32 + 16 * q, hi);
      } else {
        ycol[QY_DSDF * 4] = tf32_hi(dsdf);
        float dch[8], dcl[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) dch[i] = dcl[i] = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          dch[i] = tf32_hi(dc[i]);
          dcl[i] = dc[i] - dch[i];
          ycol[(QY_DC + i) * 4] = dch[i];
        }
        tmem_st8(lane_tb + Q_DCH, dch);              // A operand of dh3 = W4^T dc (issued with phase 2)
        tmem_st8(lane_tb + Q_DCL, dcl);
        // uncertainty grid: raw[..., 4] is the trilinear sample itself
        if (active && grads.uncert && du != 0.f) {
          const UncertPos up = uncert_pos(P, y0, y1, y2);
#pragma unroll
          for (int cnr = 0; cnr < 8; ++cnr) {
            float w;
            const int off = uncert_corner(P, up, cnr, w);
            if (off >= 0) atomicAdd(grads.uncert + off, w * du);
          }
        }
      }
      // ---- phase 1: h1 = relu(W1 [hash | oneblob]) ----
      {
        QT0;
        q_publish();
        QT1(tw_pub);
      }
      if (issuer) {
        tc_fence_after();
        if (elect_one()) {
          q_issue_1p<80, 32>(tb + Q_ACC, tb + Q_AHI, w_s + QW_W1 * 4);
          mma_commit(&bars->mma);
        }
      }
      {
        QT0;
        q_wait(&bars->mma, ph_mma);
        QT1(tw_mma);
      }
      {
        float h[8], hi[8];
        tmem_ld8(lane_tb + Q_ACC + 8 * q, h);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          hi[i] = tf32_hi(fmaxf(h[i], 0.f));
          xcol[(QX_H1 + 8 * q + i) * 4] = hi[i];
        }
        tmem_st8(lane_tb + Q_AHI + 8 * q, hi);
      }
      // ---- phase 2: a3 = W23 h1 + W3_ob oneblob ----
      {
        QT0;
        q_publish();
        QT1(tw_pub);
      }
      if (issuer) {
        tc_fence_after();
        if (elect_one()) {
          q_issue_1p<80, 32>(tb + Q_ACC, tb + Q_AHI, w_s + QW_A3 * 4);
          q_issue_3p<8, 32>(tb + Q_ACC2, tb + Q_DCH, tb + Q_DCL, w_s + QW_W4TH * 4, w_s + QW_W4TL * 4);   // dh3 = W4^T dc
          mma_commit(&bars->mma);
          // TMA: the next tile's inputs.  Its slot was last used by tile k - 2, which the scatter warps finished long ago
          if (k + 1 < my_tiles) {
            const int s1 = s + 1 == Q_NSLOT ? 0 : s + 1;
            {
              QT0;
              if (k + 1 >= Q_NSLOT) mbar_wait(&bars->free_[s1], (uint32_t)(((k + 1) / Q_NSLOT - 1) & 1));
              QT1(tw_free);
            }
            q_prefetch(slots + s1 * QS_FLOATS, &bars->full[s1], tl + gridDim.x, feat, draw, masks);
          }
        }
      }
      {
        QT0;
        q_wait(&bars->mma, ph_mma);
        QT1(tw_mma);
      }
      {
        float a3[8], dh3[8], hi[8], lo[8];
        tmem_ld8(lane_tb + Q_ACC + 8 * q, a3);
        tmem_ld8(lane_tb + Q_ACC2 + 8 * q, dh3);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int j = 8 * q + i;
          xcol[(QX_H3 + j) * 4] = tf32_hi(fmaxf(a3[i], 0.f));
          // da3 = dh3 * relu'   (dh3 = W4^T dc came off the tensor core with a3)
          const float v = (m3 >> i) & 1u ? dh3[i] : 0.f;
          hi[i] = tf32_hi(v);
          lo[i] = v - hi[i];
          ycol[(QY_DA3 + j) * 4] = hi[i];
        }
        tmem_st8(lane_tb + Q_AHI + 8 * q, hi);
        tmem_st8(lane_tb + Q_ALO + 8 * q, lo);
        if (q == 0) {                                      // [dsdf, 0 x 7] behind da3 (the first eight, now dead, OneBlob columns)
          float dh[8], dl[8];
          dh[0] = tf32_hi(dsdf);
          dl[0] = dsdf - dh[0];
#pragma unroll
          for (int i = 1; i < 8; ++i) dh[i] = dl[i] = 0.f;
          tmem_st8(lane_tb + Q_AHI + 32, dh);
          tmem_st8(lane_tb + Q_ALO + 32, dl);
        }
      }
      // ---- phase B12: dh1 = W23^T da3 + W2[0]^T dsdf ----
      {
        QT0;
        q_publish();
        QT1(tw_pub);
      }
      if (issuer) {
        tc_fence_after();
        if (elect_one()) {
          q_issue_3p<40, 32>(tb + Q_ACC, tb + Q_AHI, tb + Q_ALO, w_s + QW_B12H * 4, w_s + QW_B12L * 4);
          mma_commit(&bars->mma);
        }
      }
      {
        QT0;
        q_wait(&bars->mma, ph_mma);
        QT1(tw_mma);
      }
      {
        float dh[8], hi[8], lo[8];
        tmem_ld8(lane_tb + Q_ACC + 8 * q, dh);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float v = (m1 >> i) & 1u ? dh[i] : 0.f;
          hi[i] = tf32_hi(v);
          lo[i] = v - hi[i];
          ycol[(8 * q + i) * 4] = hi[i];
        }
        tmem_st8(lane_tb + Q_AHI + 8 * q, hi);
        tmem_st8(lane_tb + Q_ALO + 8 * q, lo);
      }
      // ---- dfeat = W1[:, :32]^T da1, and the weight-gradient GEMM D += Y^T X over this tile's 128 points ----
      fence_async_smem();          // X^T / Y^T were written with generic st.shared; the MMA reads them through the async proxy
      {
        QT0;
        q_publish();
        QT1(tw_pub);
      }
      if (issuer) {
        tc_fence_after();
        if (elect_one()) {
          q_issue_3p<32, 32>(tb + Q_ACC, tb + Q_AHI, tb + Q_ALO, w_s + QW_W1TH * 4, w_s + QW_W1TL * 4);
          mma_commit(&bars->mma);                   // dfeat: waited for right below
          constexpr uint32_t idesc = idesc_tf32(128, 144, 0, 0);
          const uint64_t yd = smem_desc(yt_s, QY_ROWS * 16, 128), xd = smem_desc(xt_s, QX_ROWS * 16, 128);
#pragma unroll
          for (int ks = 0; ks < ((dbg & 4) ? 0 : 16); ++ks)
            mma_tf32_ss(tb + Q_DW, yd + (uint64_t)(2 * QY_ROWS * ks), xd + (uint64_t)(2 * QX_ROWS * ks), idesc, k > 0 || ks > 0);
          mma_commit(&bars->wg);                    // waited for at the top of the next tile / before the flush
        }
      }
      {
        QT0;
        q_wait(&bars->mma, ph_mma);
        QT1(tw_mma);
      }
      {
        float df[8];
        tmem_ld8(lane_tb + Q_ACC + 8 * q, df);
        tmem_ld_wait();
        // the slot's feature area now takes the feature gradients, level-major [16 levels][128 rows][2] (what a scatter warp
        // reads per item is then 256 contiguous bytes).  The features were stored chunk-major, so other threads' reads of this
        // area must be over before anybody overwrites it: they are -- every MLP thread passed four group barriers since.
#pragma unroll
        for (int l = 0; l < 4; ++l) *reinterpret_cast<float2*>(slot + ((4 * q + l) * 128 + row) * 2) = make_float2(df[2 * l], df[2 * l + 1]);
      }
      if (tm == 0) s_cnt[s] = 0;                  // the slot's previous tile was drained (free) before this one was fetched
      __syncwarp();
      if ((tm & 31) == 0) mbar_arrive(&bars->ready[s]);
      // next tile
      if (++s == Q_NSLOT) {
        s = 0;
        par ^= 1u;
      }
      rem0 += step_rem;
      ray0 += step_ray;
      if (rem0 >= uS) {
        rem0 -= uS;
        ++ray0;
      }
      any_tile = true;
    }
    if (trace && tm == 0) g_q_trace[blockIdx.x * 8 + 2] = clock64();
    if (wtrace && (tm & 31) == 0) {
      long long* o = g_q_warp + (blockIdx.x * 32 + (t >> 5)) * 8;
      o[0] = clock64() - t_begin, o[1] = tw_full, o[2] = tw_mma, o[3] = tw_wg, o[4] = tw_pub, o[5] = tw_free;
    }
    if (any_tile) q_wait(&bars->wg, ph_wg);           // the last tile's weight-gradient GEMM
  }

  if (blockIdx.x == 0 && t == 0) *wg_counter = 0u;        // arms the reduction kernel's last-block counter (scratch is uninitialised)
  // ---- hand this CTA's weight-gradient block to the reduction kernel: rows 0..67 of D -> wg_part[blockIdx.x][68][144] ----
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (trace && t == 0) g_q_trace[blockIdx.x * 8 + 4] = clock64();
  if (t >= Q_SCAT) {
    const int tm = t - Q_SCAT;
    const int row = tm & 127;
    const int q = __shfl_sync(0xffffffffu, tm >> 7, 0);
    if (row < 96) {                                    // warp-uniform: rows 0..67 of D are live (lane quarters 0..2)
      const uint32_t lane_tb = tb + ((uint32_t)(32 * ((tm >> 5) & 3)) << 16);
      float* dst = wg_part + (int64_t)blockIdx.x * (QY_LIVE * 144) + row * 144;
      // nine blocks of 16 columns: quarter q takes blocks 2q and 2q + 1, quarter 0 also the ninth
#pragma unroll 1
      for (int b = 0; b < 3; ++b) {
        if (b == 2 && q != 0) break;
        const int cb = b == 2 ? 128 : 32 * q + 16 * b;
        float v[16];
        if (my_tiles > 0) {
          tmem_ld16(lane_tb + Q_DW + cb, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0.f;
        }
        if (row < QY_LIVE) {
#pragma unroll
          for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(dst + cb)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if ((t >> 5) == 0) tmem_dealloc<Q_COLS>(tb);
  if (trace && t == 0) g_q_trace[blockIdx.x * 8 + 5] = clock64();
}

// ---------------------------------------------------------------------------------------------
// Reduction of the per-CTA weight-gradient blocks (fixed order: deterministic) and the two small contractions with W2 / W3
// that stand in for the geo columns:   with  M = da3^T h1  (32 x 32)
//   dW1 = D[0:32, 0:80]      dW3[:, 0:48] = D[32:64, 32:80]      dW3[j][48+g] = sum_m M[j][m] W2[1+g][m]
//   dW2[0] = D[64, 80:112]   dW2[1+g][m] = sum_j W3[j][48+g] M[j][m]      dW4 = D[65:68, 112:144]
// Block 0 owns M and the two contractions; blocks 1..33 the 4224 elements that go straight into a gradient tensor.
// ---------------------------------------------------------------------------------------------
// Thread layout: consecutive lanes take consecutive elements (coalesced reads of every partial block); the eight 128-thread
// groups of a block take the partials c = g, g + 8, ... (all loads of a thread in flight together) and meet through shared
// memory in group order.  Blocks 0..32: the 4224 elements that go straight into a gradient tensor.  Blocks 33..40: M, 128
// elements each, into a global scratch; the last of them to finish (counter) runs the two contractions, which need all of M.
__device__ __forceinline__ float wg_sum(const float* __restrict__ part, int n_part, int e, int grp) {
  // up to 24 partial blocks per group (n_part <= 192 CTAs): every load is issued before the first add -- one trip to L2
  float v[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) {
    const int c = grp + 8 * i;
    v[i] = c < n_part ? __ldg(part + (int64_t)c * (QY_LIVE * 144) + e) : 0.f;
  }
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) acc += v[i];            // fixed order
  for (int c = grp + 192; c < n_part; c += 8) acc += __ldg(part + (int64_t)c * (QY_LIVE * 144) + e);
  return acc;
}

__device__ __forceinline__ void wg_route(int e, int& src, const NrtGrads& grads, float*& dst) {
  // direct elements: 2560 (dW1) + 1536 (dW3 oneblob part) + 32 (dW2 row 0) + 96 (dW4) = 4224
  if (e < 2560) {
    src = (e / 80) * 144 + e % 80;
    dst = grads.w1 ? grads.w1 + e : nullptr;
  } else if (e < 4096) {
    const int u = e - 2560;
    src = (32 + u / 48) * 144 + QX_OB + u % 48;
    dst = grads.w3 ? grads.w3 + (u / 48) * 63 + u % 48 : nullptr;
  } else if (e < 4128) {
    src = QY_DSDF * 144 + QX_H1 + (e - 4096);
    dst = grads.w2 ? grads.w2 + (e - 4096) : nullptr;
  } else {
    const int u = e - 4128;
    src = (QY_DC + u / 32) * 144 + QX_H3 + u % 32;
    dst = grads.w4 ? grads.w4 + u : nullptr;
  }
}

#define WG_DIRECT_BLOCKS (4224 / 128)
#define WG_M_BLOCKS 8

__global__ void __launch_bounds__(1024) wgrad_reduce_kernel(const float* __restrict__ part, int n_part, const NrtParams prm,
                                                            const NrtGrads grads, float* __restrict__ m_glob,
                                                            unsigned int* __restrict__ counter) {
  __shared__ float s_sum[8][128];
  __shared__ float Ms[32 * 33];
  __shared__ float s_w2g[NRT_GEO * 32];      // W2[1+g][m] and W3[j][48+g]: the contraction's weight operands, fetched by every M block
  __shared__ float s_w3g[32 * NRT_GEO];      // at entry so that the last one does not start two more trips to L2 behind the counter
  __shared__ unsigned int s_last;
  const int t = threadIdx.x, grp = t >> 7, el = t & 127;
  const bool m_block = blockIdx.x >= WG_DIRECT_BLOCKS;
  if (m_block) {
    if (t < NRT_GEO * 32) s_w2g[t] = __ldg(prm.w2 + 32 + t);                                   // rows 1..15 of W2, contiguous
    else if (t >= 512 && t < 512 + 32 * NRT_GEO) s_w3g[t - 512] = __ldg(prm.w3 + ((t - 512) / NRT_GEO) * 63 + NRT_OB + (t - 512) % NRT_GEO);
  }
  int e, src;
  float* dst = nullptr;
  if (m_block) {
    e = (blockIdx.x - WG_DIRECT_BLOCKS) * 128 + el;         // M = da3^T h1, element (j, m) = (e >> 5, e & 31)
    src = (32 + (e >> 5)) * 144 + QX_H1 + (e & 31);
  } else {
    e = blockIdx.x * 128 + el;
    wg_route(e, src, grads, dst);
  }
  s_sum[grp][el] = wg_sum(part, n_part, src, grp);
  __syncthreads();
  if (grp == 0) {
    float v = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) v += s_sum[g][el];
    if (m_block) m_glob[e] = v;
    else if (dst) *dst += v;
  }
  if (!m_block) return;
  // ---- the last M block contracts: dW3[j][48+g] = sum_m M[j][m] W2[1+g][m];  dW2[1+g][m] = sum_j W3[j][48+g] M[j][m] ----
  __threadfence();
  __syncthreads();
  if (t == 0) s_last = atomicAdd(counter, 1u) == WG_M_BLOCKS - 1 ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  Ms[(t >> 5) * 33 + (t & 31)] = __ldcg(m_glob + t);
  if (t == 0) *counter = 0u;                                // re-arm for the next launch
  __syncthreads();
  if (t < 480) {
    if (grads.w3) {
      const int j = t / NRT_GEO, g = t % NRT_GEO;
      float acc = 0.f;
#pragma unroll 8
      for (int m = 0; m < 32; ++m) acc = fmaf(Ms[j * 33 + m], s_w2g[g * 32 + m], acc);
      grads.w3[j * 63 + NRT_OB + g] += acc;
    }
  } else if (t >= 512 && t < 992) {
    if (grads.w2) {
      const int u = t - 512;
      const int g = u >> 5, m = u & 31;
      float acc = 0.f;
#pragma unroll 8
      for (int j = 0; j < 32; ++j) acc = fmaf(s_w3g[j * NRT_GEO + g], Ms[j * 33 + m], acc);
      grads.w2[(1 + g) * 32 + m] += acc;
    }
  }
}

size_t decode_bwd_q_smem() { return Q_SMEM_BYTES; }

int q_trace_read(void* dst, int bytes) {
  // first the [256][8] CTA stamps, then the [8][32][8] per-warp wait accounting
  const int a = (int)sizeof(long long) * 256 * 8, b = (int)sizeof(long long) * 8 * 32 * 8;
  NRT_CUDA_CHECK(cudaMemcpyFromSymbol(dst, g_q_trace, bytes < a ? bytes : a));
  if (bytes > a) NRT_CUDA_CHECK(cudaMemcpyFromSymbol((char*)dst + a, g_q_warp, bytes - a < b ? bytes - a : b));
  return NRT_OK;
}

// per-CTA weight-gradient blocks | M (1024) | counter
int64_t decode_bwd_q_scratch_floats(const NrtPlan* plan) { return (int64_t)plan->sm_count * QY_LIVE * 144 + 1024 + 32; }

int launch_decode_bwd_q(const NrtPlan* plan, const NrtParams* prm, const float* rays_o, const float* rays_d, const float* z, int S,
                        int64_t n_pts, const float* feat, const uint32_t* masks, const float* draw, const NrtGrads* grads,
                        float* wg_part, cudaStream_t st) {
  if (n_pts == 0) return NRT_OK;
  static PerDeviceOnce attr_once;
  if (attr_once.first())
    NRT_CUDA_CHECK(cudaFuncSetAttribute(decode_bwd_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Q_SMEM_BYTES));
  const int64_t tiles = (n_pts + 127) / 128;
  const int blocks = (int)(tiles < plan->sm_count ? tiles : plan->sm_count);
  static const int dbg = [] {
    const char* e = getenv("NRT_BWD_DEBUG");
    return e ? atoi(e) : 0;
  }();
  decode_bwd_q_kernel<<<blocks, Q_THREADS, Q_SMEM_BYTES, st>>>(plan->dev, *prm, rays_o, rays_d, z, S, n_pts, feat, masks, draw, *grads, wg_part,
                                                               reinterpret_cast<unsigned int*>(wg_part + (int64_t)plan->sm_count * QY_LIVE * 144 + 1024), dbg);
  NRT_CUDA_CHECK(cudaGetLastError());
  if (grads->w1 || grads->w2 || grads->w3 || grads->w4) {
    float* m_glob = wg_part + (int64_t)plan->sm_count * QY_LIVE * 144;
    wgrad_reduce_kernel<<<WG_DIRECT_BLOCKS + WG_M_BLOCKS, 1024, 0, st>>>(wg_part, blocks, *prm, *grads, m_glob,
                                                                          reinterpret_cast<unsigned int*>(m_glob + 1024));
    NRT_CUDA_CHECK(cudaGetLastError());
  }
  return NRT_OK;
}
