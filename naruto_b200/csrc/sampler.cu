// Device-resident ray sampling for the mapping iteration (SURVEY.md section 8, rows a1-a5 / "next" row f1): what the reference
// does on the host with Python lists, CPU tensors and numpy every iteration (key-frame database lookup, current-frame
// sampling, pose transform, uncertainty-aware ActiveRaySampler) as a handful of small kernels on device-resident data.
//
//   camera_rays_kernel       tp/datasets/utils.py:24-57               (pinhole directions, once per run)
//   pack_frame_kernel        src/slam/coslam/coslam.py:290-291        (direction | rgb | depth -> [H*W,7])
//   valid_depth_count_kernel src/slam/coslam/coslam.py:319-321        (number of pixels with 0 < depth <= depth_trunc)
//   kf_store_kernel          src/slam/coslam/model/keyframe.py:21-60  (rows of the frame -> one key-frame slot, doubling rule)
//   feistel_sample_kernel    random.sample(range(n), k)               (k distinct uniform indices: a keyed bijection of [0,n))
//   assemble_rays_kernel     tp/model/keyframe.py:69-79 + src/slam/coslam/coslam.py:329-344
//   pool_uncert_kernel, active_select_kernel   src/slam/coslam/active_ray_sampler.py:105-149
//
// Index lists are inputs (the reference's `random.sample` draws, for parity) or come from feistel_sample_kernel.
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// camera / frame packing
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) camera_rays_kernel(int H, int W, float fx, float fy, float cx, float cy, float* __restrict__ dirs) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (int64_t)H * W) return;
  const int i = (int)(p % W), j = (int)(p / W);
  dirs[p * 3 + 0] = __fdiv_rn(__fsub_rn((float)i, cx), fx);
  dirs[p * 3 + 1] = -__fdiv_rn(__fsub_rn((float)j, cy), fy);
  dirs[p * 3 + 2] = -1.0f;
}

__global__ void __launch_bounds__(256) pack_frame_kernel(const float* __restrict__ dir, const float* __restrict__ rgb,
                                                         const float* __restrict__ depth, int64_t n, float* __restrict__ out) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  float* o = out + p * 7;
  o[0] = dir[p * 3];
  o[1] = dir[p * 3 + 1];
  o[2] = dir[p * 3 + 2];
  o[3] = rgb[p * 3];
  o[4] = rgb[p * 3 + 1];
  o[5] = rgb[p * 3 + 2];
  o[6] = depth[p];
}

__global__ void __launch_bounds__(256) valid_depth_count_kernel(const float* __restrict__ rays, int64_t n, float depth_trunc,
                                                                int* __restrict__ count) {
  int c = 0;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const float d = rays[p * 7 + 6];
    c += (d > 0.0f && d <= depth_trunc) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

// dst[i] = frame[idxs[i % n_idx]]: the reference doubles the selected rows until there are at least P of them
// n_valid_dev (optional): device count of valid-depth pixels the indices were drawn from; when it is 0 the slot is left untouched,
// as the reference does (it attaches the frame id and returns before storing when no pixel has a valid depth)
__global__ void __launch_bounds__(256) kf_store_kernel(const float* __restrict__ frame, const int64_t* __restrict__ idxs,
                                                       int64_t n_idx, int P, const int* __restrict__ n_valid_dev,
                                                       float* __restrict__ dst) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)P * 7) return;
  if (n_valid_dev && *n_valid_dev <= 0) return;
  const int64_t i = t / 7;
  const int k = (int)(t - i * 7);
  dst[t] = frame[idxs[i % n_idx] * 7 + k];
}

// ---------------------------------------------------------------------------------------------
// k distinct uniform indices out of [0, n): the first k images of a keyed bijection of [0, n) (balanced Feistel network on
// the enclosing power of four, cycle-walked back into range).  Same distribution family as random.sample: a uniform draw
// without replacement; the stream itself is of course a different one.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}

__global__ void __launch_bounds__(256) feistel_sample_kernel(int64_t n, int64_t k, uint64_t seed, const int* __restrict__ n_dev,
                                                             int64_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k) return;
  if (n_dev) n = *n_dev;                       // population size computed on the device (valid-depth count)
  if (n <= 0) {
    out[i] = 0;
    return;
  }
  int half_bits = 1;
  while (((int64_t)1 << (2 * half_bits)) < n) ++half_bits;
  const uint64_t half_mask = ((uint64_t)1 << half_bits) - 1;
  uint64_t x = (uint64_t)(i % n);              // k > n only happens for degenerate frames; indices then repeat
  do {
    uint32_t l = (uint32_t)(x >> half_bits), r = (uint32_t)(x & half_mask);
#pragma unroll
    for (int round = 0; round < 6; ++round) {
      const uint32_t f = mix32(r ^ (uint32_t)(seed >> (8 * (round & 3))) ^ (0x9E3779B9u * (uint32_t)(round + 1))) ^ (uint32_t)(seed >> 32);
      const uint32_t nl = r;
      r = (l ^ f) & (uint32_t)half_mask;
      l = nl;
    }
    x = ((uint64_t)l << half_bits) | r;
  } while (x >= (uint64_t)n);
  out[i] = (int64_t)x;
}

// ---------------------------------------------------------------------------------------------
// gather global + current rays and move them into the world frame
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) assemble_rays_kernel(const float* __restrict__ kf_rays, const int64_t* __restrict__ frame_ids,
                                                            int P, int keyframe_every, const int64_t* __restrict__ idxs_g,
                                                            int64_t n_g, const float* __restrict__ cur_rays,
                                                            const int64_t* __restrict__ idx_cur, int64_t n_c,
                                                            const float* __restrict__ poses, int n_pose, float* __restrict__ rays_o,
                                                            float* __restrict__ rays_d, float* __restrict__ target_s,
                                                            float* __restrict__ target_d) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_g + n_c) return;
  const float* src;
  int64_t pose_row;
  if (r < n_g) {
    const int64_t idx = idxs_g[r];
    src = kf_rays + idx * 7;
    pose_row = frame_ids[idx / P] / keyframe_every;          // torch.div(ids, keyframe_every, rounding_mode='trunc')
  } else {
    src = cur_rays + idx_cur[r - n_g] * 7;
    pose_row = n_pose - 1;                                    // index -1: the current frame's pose is the last row
  }
  const float* M = poses + pose_row * 16;
  const float d0 = src[0], d1 = src[1], d2 = src[2];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    // torch.sum(d_cam[k] * R[a][k], -1): rounded products, summed left to right
    const float p0 = __fmul_rn(d0, M[a * 4 + 0]), p1 = __fmul_rn(d1, M[a * 4 + 1]), p2 = __fmul_rn(d2, M[a * 4 + 2]);
    rays_d[r * 3 + a] = __fadd_rn(__fadd_rn(p0, p1), p2);
    rays_o[r * 3 + a] = M[a * 4 + 3];
    target_s[r * 3 + a] = src[3 + a];
  }
  target_d[r] = src[6];
}

// ---------------------------------------------------------------------------------------------
// ActiveRaySampler
// ---------------------------------------------------------------------------------------------
// order-preserving map float -> uint32 (ascending)
__device__ __forceinline__ uint32_t float_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// uncertainty of every pool ray (rows base .. N - tail): voxel lookup at the back-projected end point
__global__ void __launch_bounds__(256) pool_uncert_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                          const float* __restrict__ target_d, int64_t base, int64_t n_pool,
                                                          const float* __restrict__ vol, int X, int Y, int Z, float b0, float b1,
                                                          float b2, uint32_t* __restrict__ keys) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_pool) return;
  const int64_t r = base + t;
  const float td = target_d[r];
  const float bb[3] = {b0, b1, b2};
  const int dims[3] = {X, Y, Z};
  int id[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float pt = __fadd_rn(rays_o[r * 3 + a], __fmul_rn(rays_d[r * 3 + a], td));
    const float loc = __fmul_rn(__fsub_rn(pt, bb[a]), 10.0f);
    const float rr = rintf(loc);                               // numpy round: half to even
    int v = rr >= 2147483000.0f ? 2147483000 : rr <= -2147483000.0f ? -2147483000 : (int)rr;
    id[a] = min(max(v, 0), dims[a] - 1);
  }
  keys[t] = float_key(vol[((int64_t)id[0] * Y + id[1]) * Z + id[2]]);
}

// One CTA: the K smallest keys of the pool (radix select, ties at the K-th value broken by lowest index), written in
// ascending pool order, followed by the recombination [K chosen pool rows | rows 0 .. base-K | last `tail` rows].
#define SEL_THREADS 1024
__global__ void __launch_bounds__(SEL_THREADS, 1) active_select_kernel(const uint32_t* __restrict__ keys, int64_t n_pool, int K,
                                                                       int64_t base, int64_t N, int64_t tail,
                                                                       const float* __restrict__ in_o, const float* __restrict__ in_d,
                                                                       const float* __restrict__ in_s, const float* __restrict__ in_t,
                                                                       float* __restrict__ out_o, float* __restrict__ out_d,
                                                                       float* __restrict__ out_s, float* __restrict__ out_t,
                                                                       int* __restrict__ chosen) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned int s_prefix, s_remaining, s_warp[SEL_THREADS / 32], s_base_lt, s_base_eq;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  // ---- radix select of the K-th smallest key ----
  unsigned int prefix = 0, mask = 0, remaining = (unsigned)K;       // remaining: rank (1-based) inside the current bucket
  for (int shift = 24; shift >= 0; shift -= 8) {
    if (t < 256) hist[t] = 0;
    __syncthreads();
    for (int64_t i = t; i < n_pool; i += SEL_THREADS) {
      const uint32_t k = keys[i];
      if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (t == 0) {
      unsigned int acc = 0;
      int b = 0;
      for (; b < 256; ++b) {
        if (acc + hist[b] >= remaining) break;
        acc += hist[b];
      }
      s_prefix = prefix | ((unsigned)b << shift);
      s_remaining = remaining - acc;
    }
    __syncthreads();
    prefix = s_prefix;
    remaining = s_remaining;
    mask |= 255u << shift;
    __syncthreads();
  }
  const uint32_t kth = prefix;                 // the K-th smallest key; `remaining` of the keys equal to it are taken
  // ---- ordered compaction: keys < kth, and the first `remaining` keys == kth ----
  if (t == 0) {
    s_base_lt = 0;
    s_base_eq = 0;
  }
  __syncthreads();
  for (int64_t i0 = 0; i0 < n_pool; i0 += SEL_THREADS) {
    const int64_t i = i0 + t;
    const uint32_t k = i < n_pool ? keys[i] : 0xFFFFFFFFu;
    const bool lt = i < n_pool && k < kth, eq = i < n_pool && k == kth;
    // block-wide exclusive ranks of lt and eq flags (two ballots + warp totals)
    const unsigned blt = __ballot_sync(0xffffffffu, lt), beq = __ballot_sync(0xffffffffu, eq);
    const unsigned lower = (1u << lane) - 1u;
    if (lane == 0) s_warp[warp] = (unsigned)__popc(blt) | ((unsigned)__popc(beq) << 16);
    __syncthreads();
    unsigned off_lt = 0, off_eq = 0;
    for (int w = 0; w < warp; ++w) {
      off_lt += s_warp[w] & 0xFFFFu;
      off_eq += s_warp[w] >> 16;
    }
    unsigned tot_lt = 0, tot_eq = 0;
    for (int w = 0; w < SEL_THREADS / 32; ++w) {
      tot_lt += s_warp[w] & 0xFFFFu;
      tot_eq += s_warp[w] >> 16;
    }
    const unsigned rank_lt = s_base_lt + off_lt + __popc(blt & lower);
    const unsigned rank_eq = s_base_eq + off_eq + __popc(beq & lower);
    // slot in the output: all selected rows in ascending pool order.  A row with key == kth is selected iff its rank among
    // the equal keys is < remaining; its slot = (#selected lt rows before it) + rank_eq ... computed as a merged order:
    // selected rows before i = (lt rows before i) + min(eq rows before i, remaining)
    if (lt || (eq && rank_eq < remaining)) {
      const unsigned slot = rank_lt + min(rank_eq, remaining);
      const int64_t src = base + i;
      out_o[slot * 3 + 0] = in_o[src * 3 + 0];
      out_o[slot * 3 + 1] = in_o[src * 3 + 1];
      out_o[slot * 3 + 2] = in_o[src * 3 + 2];
      out_d[slot * 3 + 0] = in_d[src * 3 + 0];
      out_d[slot * 3 + 1] = in_d[src * 3 + 1];
      out_d[slot * 3 + 2] = in_d[src * 3 + 2];
      out_s[slot * 3 + 0] = in_s[src * 3 + 0];
      out_s[slot * 3 + 1] = in_s[src * 3 + 1];
      out_s[slot * 3 + 2] = in_s[src * 3 + 2];
      out_t[slot] = in_t[src];
      if (chosen) chosen[slot] = (int)i;
    }
    __syncthreads();
    if (t == 0) {
      s_base_lt += tot_lt;
      s_base_eq += tot_eq;
    }
    __syncthreads();
  }
  // ---- the untouched parts: rows 0 .. base-K, then the last `tail` rows ----
  const int64_t keep = base - K;
  for (int64_t j = t; j < keep + tail; j += SEL_THREADS) {
    const int64_t src = j < keep ? j : N - tail + (j - keep);
    const int64_t dst = K + j;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      out_o[dst * 3 + a] = in_o[src * 3 + a];
      out_d[dst * 3 + a] = in_d[src * 3 + a];
      out_s[dst * 3 + a] = in_s[src * 3 + a];
    }
    out_t[dst] = in_t[src];
  }
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
static inline unsigned blocks_for(int64_t n, int per = 256) { return (unsigned)((n + per - 1) / per); }

int launch_camera_rays(int H, int W, float fx, float fy, float cx, float cy, float* dirs, cudaStream_t st) {
  camera_rays_kernel<<<blocks_for((int64_t)H * W), 256, 0, st>>>(H, W, fx, fy, cx, cy, dirs);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}

int launch_pack_frame(const float* dir, const float* rgb, const float* depth, int64_t n, float* out, cudaStream_t st) {
  if (n == 0) return NRT_OK;
  pack_frame_kernel<<<blocks_for(n), 256, 0, st>>>(dir, rgb, depth, n, out);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}

int launch_valid_depth_count(const float* rays, int64_t n, float depth_trunc, int* count, cudaStream_t st) {
  NRT_CUDA_CHECK(cudaMemsetAsync(count, 0, sizeof(int), st));
  if (n == 0) return NRT_OK;
  unsigned b = blocks_for(n);
  if (b > 592) b = 592;
  valid_depth_count_kernel<<<b, 256, 0, st>>>(rays, n, depth_trunc, count);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}

int launch_kf_store(const float* frame, const int64_t* idxs, int64_t n_idx, int P, const int* n_valid_dev, float* dst, cudaStream_t st) {
  kf_store_kernel<<<blocks_for((int64_t)P * 7), 256, 0, st>>>(frame, idxs, n_idx, P, n_valid_dev, dst);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}

int launch_feistel_sample(int64_t n, int64_t k, uint64_t seed, const int* n_dev, int64_t* out, cudaStream_t st) {
  if (k == 0) return NRT_OK;
  feistel_sample_kernel<<<blocks_for(k), 256, 0, st>>>(n, k, seed, n_dev, out);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}

int launch_assemble_rays(const float* kf_rays, const int64_t* frame_ids, int P, int keyframe_every, const int64_t* idxs_g,
                         int64_t n_g, const float* cur_rays, const int64_t* idx_cur, int64_t n_c, const float* poses, int n_pose,
                         float* rays_o, float* rays_d, float* target_s, float* target_d, cudaStream_t st) {
  if (n_g + n_c == 0) return NRT_OK;
  assemble_rays_kernel<<<blocks_for(n_g + n_c), 256, 0, st>>>(kf_rays, frame_ids, P, keyframe_every, idxs_g, n_g, cur_rays, idx_cur,
                                                               n_c, poses, n_pose, rays_o, rays_d, target_s, target_d);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}

int launch_active_select(const float* rays_o, const float* rays_d, const float* target_s, const float* target_d, int64_t N,
                         int64_t n_cur, const float* vol, int X, int Y, int Z, const float* bb_min, int base, int K, int mul,
                         float* out_o, float* out_d, float* out_s, float* out_t, int* chosen, void* workspace, cudaStream_t st) {
  const int64_t tail = (n_cur + mul - 1) / mul;                 // -n_cur // mul rows from the end = ceil(n_cur / mul)
  const int64_t n_pool = N - base - tail;
  NRT_REQUIRE(n_pool >= K && K >= 1 && base >= K, "active_select: the pool must hold at least K rays and base >= K");
  uint32_t* keys = reinterpret_cast<uint32_t*>(workspace);
  pool_uncert_kernel<<<blocks_for(n_pool), 256, 0, st>>>(rays_o, rays_d, target_d, base, n_pool, vol, X, Y, Z, bb_min[0], bb_min[1],
                                                         bb_min[2], keys);
  NRT_CUDA_CHECK(cudaGetLastError());
  active_select_kernel<<<1, SEL_THREADS, 0, st>>>(keys, n_pool, K, base, N, tail, rays_o, rays_d, target_s, target_d, out_o, out_d,
                                                  out_s, out_t, chosen);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}
