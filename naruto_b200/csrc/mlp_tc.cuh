// Shared pieces of the tensor-core MLP kernels (forward_tc.cu, backward_tc.cu): weight staging, the TMEM column map,
// the three-pass tf32 layer issue and the per-CTA prologue / epilogue.
//
// One CTA = 256 threads = one tile of 128 points; point r = tensor-memory lane r is shared by the thread pair
// (r, r + 128) -- warps w and w + 4 may both touch lanes [32 (w%4), +32) -- and each thread of the pair handles one half of
// the columns of every activation (kTcHalf() = threadIdx.x >> 7).  Activations are staged by their owner threads into
// TMEM as tf32 hi/lo pieces (A operand of tcgen05.mma, A-from-TMEM form); weights are the B operand in shared memory
// (chunk-major K-major, hi and lo copies).
#pragma once
#include "common.cuh"
#include "umma.cuh"

using namespace umma;

// Layer fusion.  The SDF net's second layer has no activation, so the colour net's first layer can be folded onto h1:
//     a3 = W3[:, :48] oneblob + W3[:, 48:63] geo,   geo = W2[1:16, :] h1   =>   a3 = W3[:, :48] oneblob + W23 h1,
//     W23 = W3[:, 48:63] W2[1:16, :]  (32 x 32, formed once per CTA in fp32).
// One tensor-core phase on A = [h1 | oneblob] (K = 80) therefore yields o = W2 h1 (16) and a3 (32) together (N = 48):
// three dependent phases per forward tile instead of four, and the same trick merges two backward phases.
//
// shared-memory forward weight block (floats): chunk-major K-major B operands; the lo pieces follow at +FW_FLOATS
#define FW_W1 0                    // [20 K-chunks][32 rows j][4]   k: hash 0..31 | oneblob 32..79
#define FW_W23 (FW_W1 + 80 * 32)   // [20][48 rows][4]  k: h1 0..31 | oneblob 32..79; rows 0..15 = [W2 | 0], 16..47 = [W23 | W3_ob]
#define FW_W4 (FW_W23 + 80 * 48)   // [ 8][16 rows i][4]            rows 3..15 = 0
#define FW_FLOATS (FW_W4 + 32 * 16)

// TMEM columns common to both kernels
#define TC_ACC 0                   // [0,64)    accumulator of the current phase (up to N = 48)
#define TC_AHI 64                  // [64,144)  A_hi: X0[32] (hash features, later relu(h1), relu(h3), ...) | OneBlob[48]
#define TC_ALO 144                 // [144,224) A_lo: same structure, the low-order pieces
#define TA_X0 0
#define TA_OB 32

#define TC_THREADS 256             // MLP threads per CTA (thread pairs of 128 rows)
#define TC_SMEM_HEADER 128
#define TC_SMEM_WEIGHTS (TC_SMEM_HEADER + 2 * FW_FLOATS * 4)

__device__ __forceinline__ void put_split(float* hi_blk, int o, float v) {
  const float h = tf32_hi(v);
  hi_blk[o] = h;
  hi_blk[o + FW_FLOATS] = v - h;
}

// W23[j][m] = sum_g w3[j][48+g] * w2[1+g][m]
__device__ __forceinline__ float w23_at(const NrtParams& prm, int j, int m) {
  float acc = 0.f;
#pragma unroll
  for (int g = 0; g < NRT_GEO; ++g) acc = fmaf(__ldg(prm.w3 + j * 63 + NRT_OB + g), __ldg(prm.w2 + (1 + g) * 32 + m), acc);
  return acc;
}

__device__ __forceinline__ void load_weights_tc(float* sw, const NrtParams& prm) {
  for (int i = threadIdx.x; i < 80 * 32; i += blockDim.x) {
    const int j = i / 80, k = i % 80;
    put_split(sw, FW_W1 + ((k >> 2) * 32 + j) * 4 + (k & 3), __ldg(prm.w1 + i));
  }
  for (int i = threadIdx.x; i < 48 * 80; i += blockDim.x) {
    const int n = i / 80, k = i % 80;
    float v;
    if (n < 16) v = k < 32 ? __ldg(prm.w2 + n * 32 + k) : 0.f;
    else v = k < 32 ? w23_at(prm, n - 16, k) : __ldg(prm.w3 + (n - 16) * 63 + (k - 32));
    put_split(sw, FW_W23 + ((k >> 2) * 48 + n) * 4 + (k & 3), v);
  }
  for (int i = threadIdx.x; i < 16 * 32; i += blockDim.x) {
    const int r = i >> 5, k = i & 31;
    put_split(sw, FW_W4 + ((k >> 2) * 16 + r) * 4 + (k & 3), r < 3 ? __ldg(prm.w4 + r * 32 + k) : 0.f);
  }
}

// The same weight image built from a shared-memory staging area: the raw weights come in with ONE round of global loads per
// thread (all in flight before the first store), W23 is formed once from shared memory, and the chunk-major hi / lo operands
// are written from there.  load_weights_tc() above issues ~30 dependent-latency global loads per W23 entry and per CTA; measured
// in the backward kernel, this staging cut the prologue from 12 k to 9 k cycles per CTA.  `scratch`: >= 5760 floats of shared
// memory that are free during the prologue; NT = blockDim.x (>= 672).  Contains two __syncthreads().
template <int NT>
__device__ __forceinline__ void load_weights_tc_staged(float* sw, const NrtParams& prm, float* scratch) {
  static_assert(NT >= 672, "one round of loads per thread needs at least 672 threads");
  float* raw1 = scratch;             // w1 [32][80]
  float* raw2 = raw1 + 2560;         // w2 [16][32]
  float* raw3 = raw2 + 512;          // w3 [32][63]
  float* w23s = raw3 + 2016 + 32;    // W23 [32][32]
  float* raw4 = w23s + 1024;         // w4 [3][32]
  const int t = threadIdx.x;
  {
    float a[4], c[3];
#pragma unroll
    for (int k = 0; k < 4; ++k) a[k] = t + k * NT < 2560 ? __ldg(prm.w1 + t + k * NT) : 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) c[k] = t + k * NT < 2016 ? __ldg(prm.w3 + t + k * NT) : 0.f;
    const float b = t < 512 ? __ldg(prm.w2 + t) : 0.f;
    const float d = t < 96 ? __ldg(prm.w4 + t) : 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (t + k * NT < 2560) raw1[t + k * NT] = a[k];
#pragma unroll
    for (int k = 0; k < 3; ++k)
      if (t + k * NT < 2016) raw3[t + k * NT] = c[k];
    if (t < 512) raw2[t] = b;
    if (t < 96) raw4[t] = d;
  }
  __syncthreads();
  for (int i = t; i < 1024; i += NT) {                          // W23[j][m] = sum_g w3[j][48+g] * w2[1+g][m]   (as w23_at)
    const int j = i >> 5, m = i & 31;
    float acc = 0.f;
#pragma unroll
    for (int g = 0; g < NRT_GEO; ++g) acc = fmaf(raw3[j * 63 + NRT_OB + g], raw2[(1 + g) * 32 + m], acc);
    w23s[i] = acc;
  }
  __syncthreads();
  for (int i = t; i < 80 * 32; i += NT) {
    const int j = i / 80, k = i % 80;
    put_split(sw, FW_W1 + ((k >> 2) * 32 + j) * 4 + (k & 3), raw1[i]);
  }
  for (int i = t; i < 48 * 80; i += NT) {
    const int n = i / 80, k = i % 80;
    float v;
    if (n < 16) v = k < 32 ? raw2[n * 32 + k] : 0.f;
    else v = k < 32 ? w23s[(n - 16) * 32 + k] : raw3[(n - 16) * 63 + (k - 32)];
    put_split(sw, FW_W23 + ((k >> 2) * 48 + n) * 4 + (k & 3), v);
  }
  for (int i = t; i < 16 * 32; i += NT) {
    const int r = i >> 5, k = i & 31;
    put_split(sw, FW_W4 + ((k >> 2) * 16 + r) * 4 + (k & 3), r < 3 ? raw4[r * 32 + k] : 0.f);
  }
}

struct TileCtx {
  uint32_t tb;         // TMEM base of this CTA
  uint32_t lane_tb;    // tb + (32*warp << 16): the lanes this warp may touch
  uint64_t* bar;       // MMA-complete mbarrier
  uint32_t phase;
  uint32_t w_hi, w_lo; // shared-memory addresses of the forward weight block (hi / lo)
};

// D[:, 0:N) = A[:, a_col : a_col+K) * W^T in three tf32 passes (hi*hi + lo*hi + hi*lo); one thread issues.
// wh / wl: shared-memory byte addresses of the [K/4][N][4] hi / lo weight operand.
template <int K, int N>
__device__ __forceinline__ void issue_layer(const TileCtx& c, int a_col, uint32_t wh, uint32_t wl) {
  constexpr uint32_t idesc = idesc_tf32(128, N, 0, 0);
  const uint32_t d = c.tb + TC_ACC;
  const uint32_t ah = c.tb + TC_AHI + a_col, al = c.tb + TC_ALO + a_col;
  // one descriptor per operand; a K step of 8 (two 16-byte chunks of N rows) advances the start-address field by 2*N
  const uint64_t bh = smem_desc(wh, N * 16, 128), bl = smem_desc(wl, N * 16, 128);
#pragma unroll
  for (int ks = 0; ks < K / 8; ++ks) mma_tf32_ts(d, ah + 8 * ks, bh + (uint64_t)(2 * N * ks), idesc, ks > 0);
#pragma unroll
  for (int ks = 0; ks < K / 8; ++ks) mma_tf32_ts(d, al + 8 * ks, bh + (uint64_t)(2 * N * ks), idesc, true);
#pragma unroll
  for (int ks = 0; ks < K / 8; ++ks) mma_tf32_ts(d, ah + 8 * ks, bl + (uint64_t)(2 * N * ks), idesc, true);
}

// single pass (A_hi * B_hi): plain TF32, for values that are rounded to tf32 afterwards anyway
template <int K, int N>
__device__ __forceinline__ void issue_layer_1p(const TileCtx& c, int a_col, uint32_t wh) {
  constexpr uint32_t idesc = idesc_tf32(128, N, 0, 0);
  const uint32_t d = c.tb + TC_ACC;
  const uint32_t ah = c.tb + TC_AHI + a_col;
  const uint64_t bh = smem_desc(wh, N * 16, 128);
#pragma unroll
  for (int ks = 0; ks < K / 8; ++ks) mma_tf32_ts(d, ah + 8 * ks, bh + (uint64_t)(2 * N * ks), idesc, ks > 0);
}

// The 256 MLP threads (threads 0..255 of the CTA) meet on named barrier 1, so a CTA may carry extra warps with other roles.
#define TC_BAR_MLP 1
__device__ __forceinline__ void tc_sync() { bar_sync(TC_BAR_MLP, TC_THREADS); }

// MLP threads: make this thread's TMEM stores visible, then meet at the group barrier
__device__ __forceinline__ void layer_publish() {
  tmem_st_wait();
  tc_fence_before();
  tc_sync();
}
// all threads: wait until the MMAs committed to c.bar have completed
__device__ __forceinline__ void layer_wait(TileCtx& c) {
  __syncwarp();
  mbar_wait(c.bar, c.phase);
  c.phase ^= 1u;
  tc_fence_after();
}

// first warp of the MLP group (the group is 8 consecutive warps starting at a multiple of 8)
__device__ __forceinline__ bool tc_issuer_warp() { return ((threadIdx.x >> 5) & 7) == 0; }

// publish -> one elected lane of the group's first warp issues one layer -> wait for its accumulator
template <int K, int N>
__device__ __forceinline__ void run_layer(TileCtx& c, int a_col, uint32_t wh, uint32_t wl) {
  layer_publish();
  if (tc_issuer_warp()) {                   // the warp stays converged; one elected lane issues
    tc_fence_after();
    if (elect_one()) {
      issue_layer<K, N>(c, a_col, wh, wl);
      mma_commit(c.bar);
    }
  }
  layer_wait(c);
}

template <int K, int N>
__device__ __forceinline__ void run_layer_1p(TileCtx& c, int a_col, uint32_t wh) {
  layer_publish();
  if (tc_issuer_warp()) {
    tc_fence_after();
    if (elect_one()) {
      issue_layer_1p<K, N>(c, a_col, wh);
      mma_commit(c.bar);
    }
  }
  layer_wait(c);
}

// forward weights, hi pieces of W1 and W23 only (single-pass recompute of the backward kernel): FW_W4 floats
__device__ __forceinline__ void load_weights_hi(float* sw, const NrtParams& prm) {
  for (int i = threadIdx.x; i < 80 * 32; i += blockDim.x) {
    const int j = i / 80, k = i % 80;
    sw[FW_W1 + ((k >> 2) * 32 + j) * 4 + (k & 3)] = tf32_hi(__ldg(prm.w1 + i));
  }
  for (int i = threadIdx.x; i < 48 * 80; i += blockDim.x) {
    const int n = i / 80, k = i % 80;
    float v;
    if (n < 16) v = k < 32 ? __ldg(prm.w2 + n * 32 + k) : 0.f;
    else v = k < 32 ? w23_at(prm, n - 16, k) : __ldg(prm.w3 + (n - 16) * 63 + (k - 32));
    sw[FW_W23 + ((k >> 2) * 48 + n) * 4 + (k & 3)] = tf32_hi(v);
  }
}

// stage 16 / 8 consecutive A columns (hi and lo pieces) of this thread's row
__device__ __forceinline__ void stage16(const TileCtx& c, int col, const float* v) {
  float hi[16], lo[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    hi[i] = tf32_hi(v[i]);
    lo[i] = v[i] - hi[i];
  }
  tmem_st16(c.lane_tb + TC_AHI + col, hi);
  tmem_st16(c.lane_tb + TC_ALO + col, lo);
}
__device__ __forceinline__ void stage8(const TileCtx& c, int col, const float* v) {
  float hi[8], lo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    hi[i] = tf32_hi(v[i]);
    lo[i] = v[i] - hi[i];
  }
  tmem_st8(c.lane_tb + TC_AHI + col, hi);
  tmem_st8(c.lane_tb + TC_ALO + col, lo);
}

// which half of the columns this thread owns; read through a lane-0 broadcast so the compiler knows it is warp-uniform
// (branches on it become uniform branches, level constants indexed by it live in uniform registers)
__device__ __forceinline__ void stage4(const TileCtx& c, int col, const float* v) {
  float hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    hi[i] = tf32_hi(v[i]);
    lo[i] = v[i] - hi[i];
  }
  tmem_st4(c.lane_tb + TC_AHI + col, hi);
  tmem_st4(c.lane_tb + TC_ALO + col, lo);
}

__device__ __forceinline__ int tc_half() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 7), 0); }
__device__ __forceinline__ int tc_row() { return (int)(threadIdx.x & 127); }      // point / TMEM lane of this thread

// common prologue: forward weights -> smem, barrier init, TMEM allocation.
// smem_raw: [bar 8 | slot 4 | pad to 128 | weights hi | weights lo | kernel-specific ...]
template <int NCOLS>
__device__ __forceinline__ TileCtx cta_prologue(uint8_t* smem_raw, const NrtParams& prm, float** after_weights) {
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem_raw + 8);
  float* sw = reinterpret_cast<float*>(smem_raw + TC_SMEM_HEADER);
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<NCOLS>(slot);
  }
  load_weights_tc(sw, prm);
  TileCtx c;
  c.bar = bar;
  c.phase = 0u;
  c.w_hi = smem_u32(sw);
  c.w_lo = smem_u32(sw + FW_FLOATS);
  *after_weights = sw + 2 * FW_FLOATS;
  return c;
}
// second half of the prologue, after the kernel has written its own shared-memory operands
__device__ __forceinline__ void cta_prologue_finish(uint8_t* smem_raw, TileCtx& c) {
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tb = *reinterpret_cast<uint32_t*>(smem_raw + 8);
  c.lane_tb = c.tb + ((uint32_t)(32 * ((threadIdx.x >> 5) & 3)) << 16);
}

template <int NCOLS>
__device__ __forceinline__ void cta_epilogue(const TileCtx& c) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) tmem_dealloc<NCOLS>(c.tb);
}
