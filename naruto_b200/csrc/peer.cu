// Data-parallel exchanges of the mapping iteration over NVLink peer memory, without NCCL on the step's path.
//
// The iteration has two exchange points (naruto_b200/parallel.py): the eleven loss statistics before the backward pass and the
// flat gradient bucket before Adam.  With NCCL both are separate launches on the critical path (a latency-bound 88-byte
// all-reduce and a 6.9 MB all-reduce followed by two Adam launches).  Here every rank maps the others' buffers (symmetric
// memory: the same allocation made on every GPU and exchanged once at start-up) and the exchange happens INSIDE the kernel
// that needs the data:
//
//   stats_exchange_kernel   writes this rank's statistics into every peer's pad as flag-carrying words, polls its own pad for the
//                           peers' words, sums them in rank order (identical bits on every rank) and finalizes the losses
//                           = all-reduce + loss_finalize in one 1-CTA launch.
//   adam_peers_kernel       reduce-scatter + Adam + all-gather in one pass: rank r owns a contiguous 1/N slice of the flat
//                           parameter vector; for its slice it loads the gradient from every peer's bucket (fixed rank order),
//                           clears what it read, takes the torch-Adam step on its local moments and stores the new
//                           parameters into every peer's parameter buffer.  One kernel instead of all-reduce + 2-3 Adam
//                           launches; each rank moves (N-1)/N of the bucket once in and twice out over NVLink instead of the
//                           ring / tree traffic of an all-reduce, and Adam runs on 1/N of the parameters.
//
// Cross-GPU ordering: flags are monotonically increasing step numbers written with st.release.sys after a
// __threadfence_system() and polled with ld.acquire.sys; peer data is read with ld.volatile (never through a stale L1 line).
// Every spin is bounded (trap after ~2 s) so that a missing peer becomes a launch failure, not a hang.
#include "common.cuh"

#define PEER_MAX 8
#define FLAG_STATS 0      // flag slots of a rank's pad: [slot][PEER_MAX] uint32
#define FLAG_GRADS 1
#define FLAG_DONE 2

struct PeerTable {
  int world, rank;
  float* bucket[PEER_MAX];       // gradient bucket of every rank  [total + 1]
  float* theta[PEER_MAX];        // parameter vector of every rank [total]
  double* stats_pad[PEER_MAX];   // [world][NRT_N_STATS] doubles on every rank
  uint32_t* flags[PEER_MAX];     // [3][PEER_MAX] uint32 on every rank
  float* bucket_mc;              // NVSwitch multicast addresses of bucket / theta (all ranks at once), or null
  float* theta_mc;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_volatile_f4(const float4* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
// NVLS: one load that returns the sum over every rank's copy (reduced inside the switch), one store that lands on every rank
__device__ __forceinline__ float4 multimem_ld_reduce_f4(const float4* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_f4(float4* p, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_volatile_f32(const float* p) {
  float v;
  asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

// thread `t` < world raises this rank's flag `slot` on peer t
__device__ __forceinline__ void peer_signal(const PeerTable& T, int t, int slot, uint32_t step) {
  st_release_sys(T.flags[t] + slot * PEER_MAX + T.rank, step);
}
// thread `t` < world waits until peer t has raised flag `slot` to `step` (or beyond) on this rank
__device__ __forceinline__ void peer_wait(const PeerTable& T, int t, int slot, uint32_t step) {
  const uint32_t* f = T.flags[T.rank] + slot * PEER_MAX + t;
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(f) - step) < 0) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

// ---------------------------------------------------------------------------------------------
// statistics all-reduce + loss finalisation (one CTA of 128 threads)
// ---------------------------------------------------------------------------------------------
// xchg: this rank's exchange counter, advanced here once per iteration and read by adam_peers_kernel of the same iteration.  It
// is deliberately NOT the Adam step counter: callers restore that one after a warm-up pass (CUDA-graph capture), and a flag
// value seen twice would let a rank run past a barrier.
// The statistics travel as 8-byte {32 data bits, 32-bit exchange number} words (one store each, atomic on the wire): the receiver
// polls the word itself until it carries this iteration's number, so there is no fence and no separate flag behind the data --
// one NVLink hop instead of store, fence round trip, flag (measured at N = 2: 10.7 -> see profiles/r02d_dp_stages.log).
// Pad layout on every rank: u64 [world][2 * NRT_N_STATS], row r written by rank r (word 2k = low half of statistic k, 2k+1 = high).
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(32 * PEER_MAX) stats_exchange_kernel(const PeerTable T, double* __restrict__ stats,
                                                                       unsigned int* __restrict__ xchg, float* __restrict__ losses) {
  __shared__ uint32_t s_step;
  __shared__ double s_val[PEER_MAX][NRT_N_STATS];
  const int t = threadIdx.x;
  if (t == 0) s_step = atomicAdd(xchg, 1u) + 1u;
  __syncthreads();
  const uint32_t step = s_step;
  const int p = t >> 5, k = t & 31;                // one warp per peer, one lane per 32-bit half of a statistic
  if (p < T.world) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(stats[k >> 1]);
    const uint32_t half = (k & 1) ? (uint32_t)(bits >> 32) : (uint32_t)bits;
    st_relaxed_sys_u64(reinterpret_cast<unsigned long long*>(T.stats_pad[p]) + T.rank * (2 * NRT_N_STATS) + k,
                       ((unsigned long long)step << 32) | half);
    // ... and the same lane receives the matching word of peer p
    const unsigned long long* slot = reinterpret_cast<const unsigned long long*>(T.stats_pad[T.rank]) + p * (2 * NRT_N_STATS) + k;
    unsigned long long w;
    const long long t0 = clock64();
    while ((uint32_t)((w = ld_relaxed_sys_u64(slot)) >> 32) != step) {
      if (clock64() - t0 > 4000000000LL) __trap();
    }
    const uint32_t mine = (uint32_t)w, other = __shfl_xor_sync(0xffffffffu, mine, 1);
    if (!(k & 1)) s_val[p][k >> 1] = __longlong_as_double((long long)(((unsigned long long)other << 32) | mine));
  }
  __syncthreads();
  if (t < NRT_N_STATS) {
    double v = t == NRT_STAT_UNCERT_MIN ? 3.4e38 : 0.0;
    for (int r = 0; r < T.world; ++r) {             // rank order: identical bits on every rank
      const double pv = s_val[r][t];
      v = t == NRT_STAT_UNCERT_MIN ? fmin(v, pv) : (t < NRT_N_STATS_SUM ? v + pv : v);
    }
    if (t < NRT_N_STATS_SUM || t == NRT_STAT_UNCERT_MIN) stats[t] = v;
  }
  __syncthreads();
  if (t == 0) finalize_losses(stats, losses);
}

// ---------------------------------------------------------------------------------------------
// reduce-scatter + Adam + all-gather
// ---------------------------------------------------------------------------------------------
// NRT_PEER_DEBUG=1: globaltimer (ns) at the phase boundaries of adam_peers_kernel, CTA 0 (entry, gradients-ready barrier passed,
// slices done, fence + block count done) and the last CTA (done barrier passed) -- read back with nrt_debug_read(dst, 64 | 1<<30)
__device__ unsigned long long g_peer_trace[8];
__device__ __forceinline__ void peer_stamp(int dbg, int k) {
  if (dbg && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_peer_trace[k] = t;
  }
}

struct AdamGroup {
  int64_t begin4, end4;          // in float4 units of the flat vector
  float lr, beta1, beta2, eps, wd;
  const int* step_dev;           // Adam step count of the group (already advanced for this iteration)
  int enabled;
  int keep_grad;                 // the bucket owners clear their own copies (no zero-stores over NVLink)
};

__device__ __forceinline__ float adam1(float& p, float g, float& m, float& v, const AdamGroup& G, float step_size, float bc2s) {
  if (G.wd != 0.f) g = fmaf(G.wd, p, g);
  m = m + (g - m) * (1.0f - G.beta1);
  v = v * G.beta2 + (1.0f - G.beta2) * g * g;
  const float denom = sqrtf(v) / bc2s + G.eps;
  p = p - step_size * (m / denom);
  return p;
}

// this rank's contiguous share of one parameter group: gradients from every peer's bucket, Adam on the local moments, the
// updated parameters to every peer
__device__ __forceinline__ void adam_peers_range(const PeerTable& T, float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                                 const AdamGroup& G, float step_size, float bc2s) {
  const int W = T.world, R = T.rank, t = threadIdx.x;
  const int64_t n4 = G.end4 - G.begin4;
  const int64_t lo = G.begin4 + n4 * R / W, hi = G.begin4 + n4 * (R + 1) / W;
  const bool clear = !G.keep_grad;
  // warp-major over the grid (warp w of CTA b is global warp w * gridDim + b): a slice shorter than the grid still spreads
  // evenly over every SM instead of filling the first CTAs only
  const int64_t gw = (int64_t)(t >> 5) * gridDim.x + blockIdx.x;
  for (int64_t i = lo + gw * 32 + (t & 31); i < hi; i += (int64_t)gridDim.x * blockDim.x) {
    // every load of the element is issued before the first use: ONE trip over NVLink per element, not one per peer
    float4 th = reinterpret_cast<const float4*>(T.theta[R])[i];
    float4 m = reinterpret_cast<float4*>(exp_avg)[i], v = reinterpret_cast<float4*>(exp_avg_sq)[i];
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (T.bucket_mc) {                                              // summed inside the NVSwitch, 16 bytes come back
      g = multimem_ld_reduce_f4(reinterpret_cast<const float4*>(T.bucket_mc) + i);
    } else {
      float4 gv[PEER_MAX];
#pragma unroll
      for (int p = 0; p < PEER_MAX; ++p)
        if (p < W) gv[p] = ld_volatile_f4(reinterpret_cast<const float4*>(T.bucket[p]) + i);
#pragma unroll
      for (int p = 0; p < PEER_MAX; ++p)                            // fixed order: every rank would form the same bits
        if (p < W) g.x += gv[p].x, g.y += gv[p].y, g.z += gv[p].z, g.w += gv[p].w;
    }
    if (clear) {
      const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
      if (T.bucket_mc) multimem_st_f4(reinterpret_cast<float4*>(T.bucket_mc) + i, zero);
      else
        for (int p = 0; p < W; ++p) reinterpret_cast<float4*>(T.bucket[p])[i] = zero;   // zero_grad, by the one reader of the slice
    }
    adam1(th.x, g.x, m.x, v.x, G, step_size, bc2s);
    adam1(th.y, g.y, m.y, v.y, G, step_size, bc2s);
    adam1(th.z, g.z, m.z, v.z, G, step_size, bc2s);
    adam1(th.w, g.w, m.w, v.w, G, step_size, bc2s);
    reinterpret_cast<float4*>(exp_avg)[i] = m;
    reinterpret_cast<float4*>(exp_avg_sq)[i] = v;
    if (T.theta_mc) multimem_st_f4(reinterpret_cast<float4*>(T.theta_mc) + i, th);      // all-gather of the updated slice
    else
      for (int p = 0; p < W; ++p) reinterpret_cast<float4*>(T.theta[p])[i] = th;
  }
}

__device__ __forceinline__ void adam_group_terms(const AdamGroup& G, float* out) {
  const int st = G.enabled ? *G.step_dev : 1;
  const double bc1 = 1.0 - pow((double)G.beta1, (double)st), bc2 = 1.0 - pow((double)G.beta2, (double)st);
  out[0] = (float)((double)G.lr / bc1);
  out[1] = (float)sqrt(bc2);
}

__global__ void __launch_bounds__(256, 3) adam_peers_kernel(const __grid_constant__ PeerTable T, float* __restrict__ exp_avg,
                                                         float* __restrict__ exp_avg_sq, const __grid_constant__ AdamGroup g0,
                                                         const __grid_constant__ AdamGroup g1, const __grid_constant__ AdamGroup g2,
                                                         int64_t smooth_slot, float* __restrict__ smooth_total,
                                                         const unsigned int* __restrict__ xchg, unsigned int* __restrict__ done_counter,
                                                         int dbg) {
  __shared__ float s_c[3][2];
  if (blockIdx.x == 0) peer_stamp(dbg, 0);
  const int t = threadIdx.x;
  const uint32_t step = *xchg;                  // advanced by this iteration's stats_exchange_kernel
  // ---- every rank's gradients are complete: flag exchange (CTA 0 signals, every CTA waits) ----
  if (t < T.world) {
    if (blockIdx.x == 0) peer_signal(T, t, FLAG_GRADS, step);
    peer_wait(T, t, FLAG_GRADS, step);
  }
  if (blockIdx.x == 0) peer_stamp(dbg, 1);
  if (t == 32) adam_group_terms(g0, s_c[0]);
  if (t == 64) adam_group_terms(g1, s_c[1]);
  if (t == 96) adam_group_terms(g2, s_c[2]);
  __syncthreads();
  if (g0.enabled) adam_peers_range(T, exp_avg, exp_avg_sq, g0, s_c[0][0], s_c[0][1]);
  if (g1.enabled) adam_peers_range(T, exp_avg, exp_avg_sq, g1, s_c[1][0], s_c[1][1]);
  if (g2.enabled) adam_peers_range(T, exp_avg, exp_avg_sq, g2, s_c[2][0], s_c[2][1]);
  // the smoothness loss value rides in the slot behind the gradients: every rank forms the same sum (no zeroing needed: each
  // rank's smoothness launch clears its own slot before accumulating)
  if (blockIdx.x == 0 && t == 0 && smooth_total) {
    float s = 0.f;
    for (int p = 0; p < T.world; ++p) s += ld_volatile_f32(T.bucket[p] + smooth_slot);
    *smooth_total = s;
  }
  // ---- every rank's stores have landed before anybody's next kernel reads parameters or writes gradients ----
  // (the CTA's stores are ordered before thread 0's system-scope fence by the barrier: one fence per CTA, as in a grid sync)
  __shared__ unsigned int s_last;
  __syncthreads();
  if (blockIdx.x == 0) peer_stamp(dbg, 2);
  if (t == 0) {
    __threadfence_system();
    s_last = atomicAdd(done_counter, 1u) == gridDim.x - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (blockIdx.x == 0) peer_stamp(dbg, 3);
  if (s_last) {
    peer_stamp(dbg, 4);
    if (t == 0) *done_counter = 0u;
    if (t < T.world) {
      peer_signal(T, t, FLAG_DONE, step);
      peer_wait(T, t, FLAG_DONE, step);
    }
    __syncthreads();
    peer_stamp(dbg, 5);
  }
}

int peer_trace_read(void* dst, int bytes) {
  NRT_CUDA_CHECK(cudaMemcpyFromSymbol(dst, g_peer_trace, bytes < 64 ? bytes : 64));
  return NRT_OK;
}

int launch_stats_exchange(const NrtPeerTable* tbl, double* stats, unsigned int* xchg, float* losses, cudaStream_t st) {
  PeerTable T{};
  T.world = tbl->world, T.rank = tbl->rank;
  for (int p = 0; p < tbl->world; ++p) {
    T.bucket[p] = tbl->bucket[p], T.theta[p] = tbl->theta[p];
    T.stats_pad[p] = tbl->stats_pad[p], T.flags[p] = tbl->flags[p];
  }
  stats_exchange_kernel<<<1, 32 * PEER_MAX, 0, st>>>(T, stats, xchg, losses);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}

int launch_adam_peers(const NrtPeerTable* tbl, float* exp_avg, float* exp_avg_sq, const NrtAdamGroup* groups, int n_groups,
                      int64_t smooth_slot, float* smooth_total, const unsigned int* xchg, unsigned int* done_counter, int sm_count,
                      cudaStream_t st) {
  PeerTable T{};
  T.world = tbl->world, T.rank = tbl->rank;
  for (int p = 0; p < tbl->world; ++p) {
    T.bucket[p] = tbl->bucket[p], T.theta[p] = tbl->theta[p];
    T.stats_pad[p] = tbl->stats_pad[p], T.flags[p] = tbl->flags[p];
  }
  T.bucket_mc = tbl->bucket_mc, T.theta_mc = tbl->theta_mc;
  AdamGroup g[3] = {};
  for (int i = 0; i < 3; ++i) {
    if (i < n_groups) {
      g[i].begin4 = groups[i].begin / 4, g[i].end4 = (groups[i].end + 3) / 4;      // buffers are padded to whole float4s
      g[i].lr = groups[i].lr, g[i].beta1 = groups[i].beta1, g[i].beta2 = groups[i].beta2, g[i].eps = groups[i].eps;
      g[i].wd = groups[i].weight_decay, g[i].step_dev = groups[i].step_dev, g[i].enabled = groups[i].enabled;
      g[i].keep_grad = groups[i].keep_grad;
    }
  }
  static const int dbg = [] {
    const char* e = getenv("NRT_PEER_DEBUG");
    return e ? atoi(e) : 0;
  }();
  adam_peers_kernel<<<sm_count * 3, 256, 0, st>>>(T, exp_avg, exp_avg_sq, g[0], g[1], g[2], smooth_slot, smooth_total, xchg,
                                                  done_counter, dbg);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}
