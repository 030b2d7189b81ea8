// Shared device-side definitions for the naruto_b200 kernels (sm_100a).
//
// Fixed network shape (identical at every shipped NARUTO config: configs/Replica/replica_coslam.yaml:44-63,
// configs/MP3D/mp3d_coslam.yaml:44-63): 16-level x 2-feature hash grid, OneBlob-16 on 3 dims,
// SDF net 80->32->16 and colour net 63->32->3, bias-free with ReLU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "naruto_b200.h"

#define NRT_L 16          // hash levels
#define NRT_ENC 32        // hash feature width (L * 2)
#define NRT_BINS 16       // OneBlob bins per dim
#define NRT_OB 48         // OneBlob width
#define NRT_H 32          // hidden width of both MLPs
#define NRT_O 16          // SDF net output width: sdf + 15 geo
#define NRT_GEO 15
#define NRT_IN1 80        // 32 hash + 48 oneblob
#define NRT_IN3 63        // 48 oneblob + 15 geo
#define NRT_SMAX 256      // max samples per ray the ray kernels are built for

struct __align__(16) DevLevel {       // 48 bytes: a warp reads its level's constants from shared memory with three 16-byte loads
  float scale;        // tcnn grid_scale(level)
  uint32_t res;       // tcnn grid_resolution(scale)
  uint32_t size;      // entries in this level (hashmap_size)
  uint32_t offset;    // first entry of this level in the flat table
  uint32_t hashed;    // 1: coherent prime hash, 0: dense stride walk
  uint32_t res2;      // res*res (mod 2^32)
  uint32_t magic;     // floor(2^32 / size): umulhi(v, magic) is floor(v/size) or one less, for any 32-bit v
  uint32_t agg;       // > 0: cells are coarse relative to the sample spacing: merge runs of up to 2^agg samples before the gradient RED
  uint32_t lim;       // dense levels: size - (1 + res + res2) (0 if that is negative): a cell whose base index is below lim has all
                      // eight corners inside [0, size) without uint32 wrap-around, so the % size is the identity
  uint32_t pad_[3];
};

struct DevPlan {
  DevLevel lv[NRT_L];
  float bb_min[3];
  float bb_ext[3];      // float32(bound_max) - float32(bound_min), as the reference's tensor op gives
  int ud[3];            // uncert grid dims [Nx,Ny,Nz]
  float trunc;          // training.trunc
  float sc_trunc;       // float32(sc_factor * trunc)
  float near_z, far_z, depth_trunc, range_d;
  int n_d, n_r, S;
  // torch.linspace steps, float32((end - start) / (steps - 1)) formed in fp32 like the tensor op:
  float step_u;         // uniform ladder   linspace(near, far, n_d)
  float step_r;         // range ladder     linspace(-range_d, range_d, n_r)
  float step_n;         // near/far ladder  linspace(near, far, n_r)   (rays without a depth measurement)
};

struct NrtPlan {
  NrtConfig cfg;
  DevPlan dev;
  int64_t n_grid_floats;
  int sm_count;
};

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 ldg_f2(const float2* p) { return __ldg(p); }

__device__ __forceinline__ void red_add_f2(float2* addr, float a, float b) {
  // vectorised no-return atomic: one L2 RMW for both features of an entry
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

__device__ __forceinline__ void red_add_f4(float4* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// The two x-neighbour corner entries i0 / i1 of a cell (relative to the level's first entry `base`, which is 16-byte aligned:
// level offsets are multiples of 8 entries).  Measured on B200 (tools/micro/red_throughput.cu): a no-return reduction costs
// 0.79 ns per LANE per SM whether the lane carries 4, 8 or 16 bytes, and the backward scatter is bound by exactly that.  When
// the two entries are the halves of one aligned 16-byte pair -- dense levels with an even base index, and every hashed level
// at even x, because tcnn's coherent hash multiplies x by 1 so x -> x + 1 only flips the lowest index bit -- one v4 reduction
// carries both: 6 instead of 8 lanes per point and level on average.
__device__ __forceinline__ void red_add_xpair(float2* base, uint32_t i0, uint32_t i1, float a0, float a1, float b0, float b1) {
  if ((i0 ^ i1) == 1u) {
    const bool sw = (i0 & 1u) != 0u;
    red_add_f4(reinterpret_cast<float4*>(base + (i0 & ~1u)), sw ? b0 : a0, sw ? b1 : a1, sw ? a0 : b0, sw ? a1 : b1);
  } else {
    red_add_f2(base + i0, a0, a1);
    red_add_f2(base + i1, b0, b1);
  }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float softplusf_(float x) {
  // torch.nn.Softplus(beta=1, threshold=20)
  return x > 20.0f ? x : log1pf(expf(x));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_min_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Philox4x32-10 (counter-based; the in-kernel stand-in for the reference's torch.rand draw)
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}
__device__ __forceinline__ float u32_to_unit(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// ---------------------------------------------------------------------------------------------
// hash grid level addressing (tcnn grid.h: pos_fract + grid_index + coherent_prime_hash)
// ---------------------------------------------------------------------------------------------
struct LevelPos {
  uint32_t g[3];
  float f[3];
};

__device__ __forceinline__ LevelPos level_pos(const DevLevel& L, float x0, float x1, float x2) {
  LevelPos p;
  float xs[3] = {x0, x1, x2};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    float pos = fmaf(L.scale, xs[d], 0.5f);
    float fl = floorf(pos);
    p.g[d] = (uint32_t)(int)fl;
    p.f[d] = pos - fl;
  }
  return p;
}

// v mod L.size for any 32-bit v without a hardware divide: q = umulhi(v, floor(2^32/size)) is floor(v/size) or one
// less, so the remainder lands in [0, 2*size) and one conditional subtract finishes it
__device__ __forceinline__ uint32_t mod_size(const DevLevel& L, uint32_t v) {
  uint32_t r = v - __umulhi(v, L.magic) * L.size;
  return r >= L.size ? r - L.size : r;
}

__device__ __forceinline__ uint32_t level_index(const DevLevel& L, uint32_t cx, uint32_t cy, uint32_t cz) {
  if (L.hashed) {
    uint32_t h = cx ^ (cy * 2654435761u) ^ (cz * 805459861u);
    return h & (L.size - 1u);            // hashed levels always hold exactly 2^T entries
  }
  // dense stride walk in wrapping uint32 arithmetic (points outside the bound give "negative" coordinates), then % size
  return mod_size(L, cx + cy * L.res + cz * L.res2);
}

// trilinear weight of corner c (bit d of c selects the +1 vertex along dim d), multiplied in dim order
__device__ __forceinline__ float corner_weight(const LevelPos& p, int c) {
  float w = (c & 1) ? p.f[0] : 1.0f - p.f[0];
  w *= (c & 2) ? p.f[1] : 1.0f - p.f[1];
  w *= (c & 4) ? p.f[2] : 1.0f - p.f[2];
  return w;
}

// all eight corner entries and weights of one level; one (warp-uniform) branch per level, none per corner
__device__ __forceinline__ LevelPos level_corners(const DevLevel& L, float x0, float x1, float x2, uint32_t* __restrict__ idx,
                                                  float* __restrict__ w) {
  const LevelPos p = level_pos(L, x0, x1, x2);
  const float wx[2] = {1.0f - p.f[0], p.f[0]}, wy[2] = {1.0f - p.f[1], p.f[1]}, wz[2] = {1.0f - p.f[2], p.f[2]};
  float wxy[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) wxy[c] = wx[c & 1] * wy[c >> 1];
#pragma unroll
  for (int c = 0; c < 8; ++c) w[c] = wxy[c & 3] * wz[c >> 2];
  if (L.hashed) {
    const uint32_t hy0 = p.g[1] * 2654435761u, hz0 = p.g[2] * 805459861u;
    const uint32_t hy[2] = {hy0, hy0 + 2654435761u}, hz[2] = {hz0, hz0 + 805459861u};
    const uint32_t m = L.size - 1u;
    uint32_t hyz[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) hyz[c] = hy[c & 1] ^ hz[c >> 1];
#pragma unroll
    for (int c = 0; c < 8; ++c) idx[c] = ((p.g[0] + (c & 1)) ^ hyz[c >> 1]) & m;
  } else {
    const uint32_t base = p.g[0] + p.g[1] * L.res + p.g[2] * L.res2;
    if (base < L.lim) {                   // whole cell in range: % size is the identity (DevLevel::lim)
#pragma unroll
      for (int c = 0; c < 8; ++c) idx[c] = base + (c & 1) + ((c >> 1) & 1) * L.res + (c >> 2) * L.res2;
    } else {
#pragma unroll
      for (int c = 0; c < 8; ++c) idx[c] = mod_size(L, base + (c & 1) + ((c >> 1) & 1) * L.res + (c >> 2) * L.res2);
    }
  }
  return p;
}

// The four (y, z) corner entries of one level at x-corner xc (0 / 1), relative to the level's first entry, and the cell
// fractions.  Same arithmetic as level_corners().
__device__ __forceinline__ void level_offsets_x(const DevLevel& L, uint32_t xc, float x0, float x1, float x2,
                                                uint32_t* __restrict__ idx, float* __restrict__ fr) {
  const LevelPos p = level_pos(L, x0, x1, x2);
  fr[0] = p.f[0];
  fr[1] = p.f[1];
  fr[2] = p.f[2];
  const uint32_t gx = p.g[0] + xc;
  if (L.hashed) {
    const uint32_t hy0 = p.g[1] * 2654435761u, hz0 = p.g[2] * 805459861u;
    const uint32_t hy[2] = {hy0, hy0 + 2654435761u}, hz[2] = {hz0, hz0 + 805459861u};
    const uint32_t m = L.size - 1u;
#pragma unroll
    for (int k = 0; k < 4; ++k) idx[k] = (gx ^ hy[k & 1] ^ hz[k >> 1]) & m;
  } else {
    const uint32_t base = p.g[0] + p.g[1] * L.res + p.g[2] * L.res2;
    if (base < L.lim) {
#pragma unroll
      for (int k = 0; k < 4; ++k) idx[k] = base + xc + (k & 1) * L.res + (k >> 1) * L.res2;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) idx[k] = mod_size(L, base + xc + (k & 1) * L.res + (k >> 1) * L.res2);
    }
  }
}

// Pair-cooperative gather of NL consecutive levels.  The lanes (2i, 2i+1) of a warp serve their two points together: for
// each point the even lane reads the four x-corner-0 entries and the odd lane the four x-corner-1 entries of every level,
// so one load instruction carries 16 points x 2 x-neighbours.  x-neighbours are adjacent in dense levels and differ only in
// the low index bits under tcnn's coherent hash (prime 1 on x), i.e. they share a 32-byte sector three times out of four and
// a 128-byte line 15 times out of 16: the L1 tag look-ups and wavefronts per gathered byte -- the resource that bounds this
// kernel -- drop by ~45% against one-point-per-lane.  Each lane weights what it read (it knows both points), and one shuffle
// per feature folds the two x-halves.  Summation order: x-corner halves first (y, z order inside), then their sum.
template <int NL>
__device__ __forceinline__ void gather_levels_paired(const DevLevel* __restrict__ lv, const float2* __restrict__ grid, float x0,
                                                     float x1, float x2, float* __restrict__ f) {
  const bool odd = threadIdx.x & 1;
  const float y0 = __shfl_xor_sync(0xffffffffu, x0, 1), y1 = __shfl_xor_sync(0xffffffffu, x1, 1),
              y2 = __shfl_xor_sync(0xffffffffu, x2, 1);
  // point A belongs to the even lane of the pair, point B to the odd lane
  const float a0 = odd ? y0 : x0, a1 = odd ? y1 : x1, a2 = odd ? y2 : x2;
  const float b0 = odd ? x0 : y0, b1 = odd ? x1 : y1, b2 = odd ? x2 : y2;
  const uint32_t xc = odd ? 1u : 0u;
  uint32_t idx[NL][8];
  float fr[NL][6];
#pragma unroll
  for (int l = 0; l < NL; ++l) {
    level_offsets_x(lv[l], xc, a0, a1, a2, idx[l], fr[l]);
    level_offsets_x(lv[l], xc, b0, b1, b2, idx[l] + 4, fr[l] + 3);
  }
  float2 v[NL][8];
#pragma unroll
  for (int l = 0; l < NL; ++l) {
    const float2* base = grid + lv[l].offset;
    asm("" : "+l"(base));      // keep the level's base pointer whole: each address is then one 32x32+64 multiply-add
#pragma unroll
    for (int c = 0; c < 8; ++c) v[l][c] = ldg_f2(base + idx[l][c]);
  }
#pragma unroll
  for (int l = 0; l < NL; ++l) {
    float part[2][2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {          // q = 0: point A, 1: point B
      const float fx = fr[l][3 * q], fy = fr[l][3 * q + 1], fz = fr[l][3 * q + 2];
      const float wx = odd ? fx : 1.0f - fx;
      const float wy[2] = {1.0f - fy, fy}, wz[2] = {1.0f - fz, fz};
      float r0 = 0.f, r1 = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float w = (wx * wy[k & 1]) * wz[k >> 1];
        r0 = fmaf(w, v[l][4 * q + k].x, r0);
        r1 = fmaf(w, v[l][4 * q + k].y, r1);
      }
      part[q][0] = r0;
      part[q][1] = r1;
    }
    // the even lane keeps point A and needs the odd lane's half of it; the odd lane keeps B
    const float s0 = odd ? part[0][0] : part[1][0], s1 = odd ? part[0][1] : part[1][1];
    const float g0 = __shfl_xor_sync(0xffffffffu, s0, 1), g1 = __shfl_xor_sync(0xffffffffu, s1, 1);
    const float k0 = odd ? part[1][0] : part[0][0], k1 = odd ? part[1][1] : part[0][1];
    f[2 * l] = odd ? g0 + k0 : k0 + g0;        // x-corner-0 half first, on both lanes
    f[2 * l + 1] = odd ? g1 + k1 : k1 + g1;
  }
}

// gathers one level: returns the two interpolated features.  Summation order as gather_levels_paired(): the four (y, z)
// corners of each x-corner half, then the sum of the two halves -- every kernel of the library produces the same bits.
__device__ __forceinline__ float2 level_gather(const DevLevel& L, const float2* __restrict__ grid, float x0, float x1,
                                               float x2) {
  uint32_t idx[8];
  float w_unused[8];
  const LevelPos p = level_corners(L, x0, x1, x2, idx, w_unused);
  const float2* base = grid + L.offset;
  float2 v[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) v[c] = ldg_f2(base + idx[c]);
  const float wy[2] = {1.0f - p.f[1], p.f[1]}, wz[2] = {1.0f - p.f[2], p.f[2]};
  float2 h[2];
#pragma unroll
  for (int xc = 0; xc < 2; ++xc) {
    const float wx = xc ? p.f[0] : 1.0f - p.f[0];
    float r0 = 0.f, r1 = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float w = (wx * wy[k & 1]) * wz[k >> 1];
      r0 = fmaf(w, v[xc + 2 * k].x, r0);
      r1 = fmaf(w, v[xc + 2 * k].y, r1);
    }
    h[xc] = make_float2(r0, r1);
  }
  return make_float2(h[0].x + h[1].x, h[0].y + h[1].y);
}

// ---------------------------------------------------------------------------------------------
// uncertainty grid: torch grid_sample(vol[1,1,Nx,Ny,Nz], 2x-1, align_corners=False, zeros), with the
// reference's un-permuted coordinate order (x -> Nz axis, y -> Ny, z -> Nx)
// ---------------------------------------------------------------------------------------------
struct UncertPos {
  int i[3];      // floor index along the axis addressed by x, y, z (sizes Nz, Ny, Nx)
  float f[3];
};

__device__ __forceinline__ UncertPos uncert_pos(const DevPlan& P, float x0, float x1, float x2) {
  UncertPos u;
  float xs[3] = {x0, x1, x2};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    float g = __fsub_rn(__fmul_rn(xs[d], 2.0f), 1.0f);
    float size = (float)P.ud[2 - d];
    float c = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(g, 1.0f), size), 1.0f), 0.5f);
    float fl = floorf(c);
    u.i[d] = (int)fl;
    u.f[d] = c - fl;
  }
  return u;
}

// linear offset of corner c, or -1 when the corner is outside the volume (zero padding)
__device__ __forceinline__ int uncert_corner(const DevPlan& P, const UncertPos& u, int c, float& w) {
  int iz = u.i[0] + (c & 1), iy = u.i[1] + ((c >> 1) & 1), ix = u.i[2] + ((c >> 2) & 1);
  w = ((c & 1) ? u.f[0] : 1.0f - u.f[0]) * ((c & 2) ? u.f[1] : 1.0f - u.f[1]) * ((c & 4) ? u.f[2] : 1.0f - u.f[2]);
  bool ok = (unsigned)iz < (unsigned)P.ud[2] && (unsigned)iy < (unsigned)P.ud[1] && (unsigned)ix < (unsigned)P.ud[0];
  return ok ? (ix * P.ud[1] + iy) * P.ud[2] + iz : -1;
}

__device__ __forceinline__ float uncert_sample(const DevPlan& P, const float* __restrict__ ug, float x0, float x1,
                                               float x2) {
  UncertPos u = uncert_pos(P, x0, x1, x2);
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float w;
    int off = uncert_corner(P, u, c, w);
    float v = off >= 0 ? __ldg(ug + off) : 0.f;
    acc = fmaf(w, v, acc);
  }
  return acc;
}

// ---------------------------------------------------------------------------------------------
// OneBlob (tcnn oneblob.h): 16 bins of one coordinate
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float quartic_cdf(float t) {
  float u = t * (float)NRT_BINS;
  float u2 = u * u;
  float u4 = u2 * u2;
  float v = (15.0f / 16.0f) * u * (1.0f - (float)(2.0 / 3.0) * u2 + (float)(1.0 / 5.0) * u4) + 0.5f;
  return fminf(fmaxf(v, 0.0f), 1.0f);
}
__device__ __forceinline__ float quartic_pdf(float t) {   // d quartic_cdf / dt
  float u = t * (float)NRT_BINS;
  float u2 = u * u;
  if (u2 > 1.0f) return 0.0f;
  float m = 1.0f - u2;
  return (15.0f / 16.0f) * m * m * (float)NRT_BINS;
}
__device__ __forceinline__ float wrapped_cdf(float t) { return quartic_cdf(t) + quartic_cdf(t - 1.0f) + quartic_cdf(t + 1.0f); }

__device__ __forceinline__ void oneblob16(float x, float* __restrict__ bins) {
  float left = wrapped_cdf(0.0f - x);
  const float first = left;
#pragma unroll
  for (int b = 0; b < NRT_BINS; ++b) {
    float right = (b == NRT_BINS - 1) ? first + 1.0f : wrapped_cdf((float)(b + 1) * (1.0f / NRT_BINS) - x);
    bins[b] = right - left;
    left = right;
  }
}

// Same encoding, evaluated sparsely.  With boundary values B_b = wrapped_cdf(b/16 - x) (b = 0..15) and B_16 = B_0 + 1,
// bins[b] = B_{b+1} - B_b.  For x in [0, 1) and i0 = floor(16 x), the quartic kernel (support |u| < 1, clamped outside)
// makes every boundary except b = i0 and i0 + 1 (and, through the wrap-around, b = 0 / 16 when i0 is 0 or 15) saturate to
// exactly 0 + 0 + 1 (b < i0) or 1 + 0 + 1 (b > i0 + 1).  Of wrapped_cdf's three terms only the centre one is live at the two
// live boundaries (the -1 image is exactly 0, the +1 image exactly 1), so
//     B_i0 = qa + 1,  B_{i0+1} = qb + 1,      qa = quartic_cdf(i0/16 - x),  qb = quartic_cdf((i0+1)/16 - x)
// and only three bins are non-zero, at (i0 - 1) mod 16, i0, (i0 + 1) mod 16:
//     ca - 1 | cb - ca | 2 - cb          (i0 = 0:  first = (ca + 1) - 2 in bin 15;   i0 = 15:  last = 1 - qb in bin 0)
// -- the same expressions, in the same order, as oneblob16() evaluates for those bins, so the result equals it bit for bit
// except when x sits within an ulp of a bin boundary (difference <= 6e-8).  The three values are then rotated into place
// with a 4-stage barrel rotation whose known-zero lanes fold away at compile time (36 selects instead of a 16 x 3 compare
// chain).  x outside [0, 1) (sample points outside the bound) takes the dense path.
__device__ __forceinline__ void oneblob16_fast(float x, float* __restrict__ bins) {
  if (!(x >= 0.0f && x < 1.0f)) {
    oneblob16(x, bins);
    return;
  }
  const int i0 = min((int)(x * (float)NRT_BINS), NRT_BINS - 1);
  const float qa = quartic_cdf((float)i0 * (1.0f / NRT_BINS) - x);
  const float qb = quartic_cdf((float)(i0 + 1) * (1.0f / NRT_BINS) - x);
  const float ca = qa + 1.0f, cb = qb + 1.0f;
  const float va = i0 == 0 ? (ca + 1.0f) - 2.0f : ca - 1.0f;
  const float vb = cb - ca;
  const float vc = i0 == NRT_BINS - 1 ? 1.0f - qb : 2.0f - cb;
  // rotate [va, vb, vc, 0, ...] right by r = (i0 - 1) mod 16
  const int r = (i0 + NRT_BINS - 1) & (NRT_BINS - 1);
  float a[NRT_BINS], b[NRT_BINS];
#pragma unroll
  for (int k = 0; k < NRT_BINS; ++k) a[k] = 0.f;
  a[0] = va;
  a[1] = vb;
  a[2] = vc;
#pragma unroll
  for (int sft = 0; sft < 4; ++sft) {
    const bool on = (r >> sft) & 1;
#pragma unroll
    for (int k = 0; k < NRT_BINS; ++k) b[k] = on ? a[(k - (1 << sft)) & (NRT_BINS - 1)] : a[k];
#pragma unroll
    for (int k = 0; k < NRT_BINS; ++k) a[k] = b[k];
  }
#pragma unroll
  for (int k = 0; k < NRT_BINS; ++k) bins[k] = a[k];
}

// normalisation to the bound (tp/model/scene_rep.py:172-173): two roundings, like the tensor ops
__device__ __forceinline__ float normalise1(const DevPlan& P, int d, float p) {
  return __fdiv_rn(__fsub_rn(p, P.bb_min[d]), P.bb_ext[d]);
}

// ---------------------------------------------------------------------------------------------
// per-point decode result (the MLPs themselves run on the tensor cores: mlp_tc.cuh / forward_tc.cu)
// ---------------------------------------------------------------------------------------------
struct PointOut {
  float rgb[3];     // colour logits (half 0)
  float o8[8];      // this half's slice of the SDF-net output o[16] = [sdf, geo[15]]
  float unc;        // raw uncertainty sample (half 1)
};

// ---------------------------------------------------------------------------------------------
// depth sampling for one ray, executed by one warp (src/slam/coslam/model/scene_rep.py:158-180).
// z (smem, length S) receives the sorted (and optionally jittered) depths.
// torch.linspace semantics: step=(end-start)/(steps-1); i < steps/2 ? start+step*i : end-step*(steps-1-i)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float linspace_at(float start, float end, float step, int steps, int i) {
  if (steps == 1) return start;
  return i < steps / 2 ? __fadd_rn(start, __fmul_rn(step, (float)i)) : __fsub_rn(end, __fmul_rn(step, (float)(steps - 1 - i)));
}

__device__ __forceinline__ void warp_sample_z(const DevPlan& P, float td, const float* __restrict__ u_row, int perturb,
                                              uint64_t seed, int64_t ray, float* __restrict__ z, int lane) {
  const int nd = P.n_d, nr = P.n_r, S = P.S;
  const bool invalid = td <= 0.0f;       // reference: z_samples[target_d <= 0] = linspace(near, far)
  // near-surface ladder value j (lane j holds it; n_r <= 32 is the common case, larger ladders recompute)
  auto near_val = [&](int j) -> float {
    return invalid ? linspace_at(P.near_z, P.far_z, P.step_n, nr, j)
                   : __fadd_rn(linspace_at(-P.range_d, P.range_d, P.step_r, nr, j), td);
  };
  const float my_near = lane < nr ? near_val(lane) : 3.4e38f;
  // stable merge by rank: uniform ladder first on ties
  for (int i = lane; i < nd + ((32 - nd % 32) % 32); i += 32) {       // all lanes iterate together (shuffles inside)
    const float v = i < nd ? linspace_at(P.near_z, P.far_z, P.step_u, nd, i) : 0.f;
    int below = 0;
    if (nr <= 32) {
      for (int j = 0; j < nr; ++j) below += __shfl_sync(0xffffffffu, my_near, j) < v ? 1 : 0;
    } else {
      for (int j = 0; j < nr; ++j) below += near_val(j) < v ? 1 : 0;
    }
    if (i < nd) z[i + below] = v;
  }
  for (int j = lane; j < nr; j += 32) {
    const float v = near_val(j);
    int le = 0;
    if (nd > 0) {
      // count uniform samples <= v: guess from the spacing, then correct with the exact ladder values
      const float step = nd > 1 ? P.step_u : 1.0f;
      int g = (int)floorf((v - P.near_z) / step) + 1;
      g = max(0, min(nd, g));
      while (g < nd && linspace_at(P.near_z, P.far_z, P.step_u, nd, g) <= v) ++g;
      while (g > 0 && linspace_at(P.near_z, P.far_z, P.step_u, nd, g - 1) > v) --g;
      le = g;
    }
    z[j + le] = v;
  }
  __syncwarp();
  if (perturb) {
    // stratified jitter between neighbouring mid-points; read all neighbours before anyone writes
    float lo[NRT_SMAX / 32], hi[NRT_SMAX / 32];
#pragma unroll
    for (int p = 0; p < NRT_SMAX / 32; ++p) {
      if (p * 32 >= S) break;                 // warp-uniform: rays have S <= 256 samples, the loop is unrolled for 256
      int s = p * 32 + lane;
      if (s < S) {
        float zc = z[s];
        lo[p] = s > 0 ? 0.5f * (zc + z[s - 1]) : zc;
        hi[p] = s < S - 1 ? 0.5f * (z[s + 1] + zc) : zc;
      }
    }
    __syncwarp();
    if (!u_row) {
      // in-kernel draw: sample s takes component s & 3 of Philox(counter = (ray, s >> 2)).  One call yields four samples, so lane
      // g evaluates the call of group g once and parks the four uniforms in z[] (free now: every lane holds its lo / hi in
      // registers) instead of four lanes evaluating the same call to pick one component each.
      for (int gi = lane; 4 * gi < S; gi += 32) {
        const uint4 ctr = make_uint4((uint32_t)ray, (uint32_t)(ray >> 32), (uint32_t)gi, 0u);
        const uint4 rnd = philox4x32(ctr, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        if (4 * gi + 0 < S) z[4 * gi + 0] = u32_to_unit(rnd.x);
        if (4 * gi + 1 < S) z[4 * gi + 1] = u32_to_unit(rnd.y);
        if (4 * gi + 2 < S) z[4 * gi + 2] = u32_to_unit(rnd.z);
        if (4 * gi + 3 < S) z[4 * gi + 3] = u32_to_unit(rnd.w);
      }
      __syncwarp();
    }
#pragma unroll
    for (int p = 0; p < NRT_SMAX / 32; ++p) {
      if (p * 32 >= S) break;                 // warp-uniform: rays have S <= 256 samples, the loop is unrolled for 256
      int s = p * 32 + lane;
      if (s < S) {
        const float r = u_row ? u_row[s] : z[s];
        z[s] = __fadd_rn(lo[p], __fmul_rn(__fsub_rn(hi[p], lo[p]), r));
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// per-ray compositing by one warp (tp/model/scene_rep.py:64-84 + src/slam/coslam/model/scene_rep.py:66-96).
// raw: smem [S][5], z: smem [S].  Every lane returns the same RayOut.  Lane l owns samples l, l+32, ...; the bell weight of
// each owned sample is evaluated once and kept in registers, and the colour / uncertainty activations are only evaluated
// for samples inside the truncation window (weight > 0).
// ---------------------------------------------------------------------------------------------
struct RayOut {
  float rgb[3], depth, depth_var, acc, disp, uncert;
  float wsum;        // sum of masked bell weights (before the 1e-8)
  float z_cut;       // z_surf + sc*trunc
};

// sigma(a) * sigma(-a), a = sdf / trunc; even in a, evaluated as e / (1 + e)^2 with e = exp(-|a|) (no overflow)
__device__ __forceinline__ float bell_weight(float sdf, float trunc) {
  const float a = __fdiv_rn(sdf, trunc);
  const float e = expf(-fabsf(a));
  const float d = 1.0f + e;
  return __fdiv_rn(e, d * d);
}

__device__ __forceinline__ RayOut warp_composite(const DevPlan& P, const int S, const float* __restrict__ raw,
                                                 const float* __restrict__ z, float* __restrict__ w_out, int lane) {
  constexpr int NP = NRT_SMAX / 32;
  // first sign change (argmax of the 0/1 mask -> 0 when there is none)
  int first = 0x7fffffff;
  for (int s = lane; s < S - 1; s += 32) {
    if (raw[(s + 1) * 5 + 3] * raw[s * 5 + 3] < 0.0f) {
      first = s;
      break;
    }
  }
  first = warp_min_i(first);
  if (first == 0x7fffffff) first = 0;
  RayOut r;
  r.z_cut = __fadd_rn(z[first], P.sc_trunc);
  float bw[NP], zv[NP];
  float ws = 0.f;
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    if (p * 32 >= S) break;                   // warp-uniform early exit (the loop is unrolled for NRT_SMAX samples)
    const int s = p * 32 + lane;
    bw[p] = 0.f;
    zv[p] = 0.f;
    if (s < S) {
      zv[p] = z[s];
      if (zv[p] < r.z_cut) bw[p] = bell_weight(raw[s * 5 + 3], P.trunc);
      ws += bw[p];
    }
  }
  ws = warp_sum(ws);
  r.wsum = ws;
  const float denom = ws + 1e-8f;
  float c0 = 0, c1 = 0, c2 = 0, dep = 0, acc = 0, unc = 0;
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    if (p * 32 >= S) break;                   // warp-uniform early exit (the loop is unrolled for NRT_SMAX samples)
    const int s = p * 32 + lane;
    if (s < S) {
      const float w = zv[p] < r.z_cut ? __fdiv_rn(bw[p], denom) : 0.f;
      bw[p] = w;
      if (w_out) w_out[s] = w;
      if (w != 0.f) {
        c0 = fmaf(w, sigmoidf_(raw[s * 5 + 0]), c0);
        c1 = fmaf(w, sigmoidf_(raw[s * 5 + 1]), c1);
        c2 = fmaf(w, sigmoidf_(raw[s * 5 + 2]), c2);
        dep = fmaf(w, zv[p], dep);
        acc += w;
        unc = fmaf(w * w, softplusf_(raw[s * 5 + 4]) + 0.01f, unc);
      }
    }
  }
  r.rgb[0] = warp_sum(c0);
  r.rgb[1] = warp_sum(c1);
  r.rgb[2] = warp_sum(c2);
  r.depth = warp_sum(dep);
  r.acc = warp_sum(acc);
  r.uncert = warp_sum(unc);
  float var = 0.f;
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    if (p * 32 >= S) break;                   // warp-uniform early exit (the loop is unrolled for NRT_SMAX samples)
    const int s = p * 32 + lane;
    if (s < S) {
      const float dz = zv[p] - r.depth;
      var = fmaf(bw[p], dz * dz, var);
    }
  }
  r.depth_var = warp_sum(var);
  r.disp = 1.0f / fmaxf(1e-10f, __fdiv_rn(r.depth, r.acc));
  return r;
}

// The weight half of warp_composite() alone -- first sign change, truncation cut, normalised bell weights -> w_out[S] -- for a caller
// that already holds the ray's composited outputs (the backward pass reads them from the forward's output buffers instead of
// integrating colour / depth / uncertainty a second time).  Same expressions, same bits as warp_composite().
__device__ __forceinline__ void warp_weights(const DevPlan& P, const int S, const float* __restrict__ raw, const float* __restrict__ z,
                                             float* __restrict__ w_out, int lane, float& z_cut, float& wsum) {
  constexpr int NP = NRT_SMAX / 32;
  int first = 0x7fffffff;
  for (int s = lane; s < S - 1; s += 32) {
    if (raw[(s + 1) * 5 + 3] * raw[s * 5 + 3] < 0.0f) {
      first = s;
      break;
    }
  }
  first = warp_min_i(first);
  if (first == 0x7fffffff) first = 0;
  z_cut = __fadd_rn(z[first], P.sc_trunc);
  float bw[NP];
  float ws = 0.f;
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    if (p * 32 >= S) break;
    const int s = p * 32 + lane;
    bw[p] = 0.f;
    if (s < S) {
      if (z[s] < z_cut) bw[p] = bell_weight(raw[s * 5 + 3], P.trunc);
      ws += bw[p];
    }
  }
  ws = warp_sum(ws);
  wsum = ws;
  const float denom = ws + 1e-8f;
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    if (p * 32 >= S) break;
    const int s = p * 32 + lane;
    if (s < S) w_out[s] = z[s] < z_cut ? __fdiv_rn(bw[p], denom) : 0.f;
  }
}

// Saved hash features (NrtRenderOut::feat) are stored tile-major: point p, chunk c (4 floats = the two features of two levels)
// lives at float4 index ((p / 128) * 8 + c) * 128 + p % 128, i.e. [tile of 128 points][8 chunks][128 points][4 floats].  Both the
// writers (gather threads: one chunk of 32 consecutive points) and the readers (backward thread pairs: chunks of 32 consecutive
// points) then touch 512 contiguous bytes per warp instruction instead of 32 scattered 16-byte pieces.
__device__ __forceinline__ int64_t feat_tiled_index(int64_t p, int chunk) { return ((p >> 7) * 8 + chunk) * 128 + (p & 127); }

// where the points of a launch come from: an explicit [n,3] array, or rays + depths
struct PointSource {
  const float* x;                 // [n,3] normalised points, or NULL -> rays
  const float* rays_o;
  const float* rays_d;
  const float* z;                 // [B,S]
  int S;
};

__device__ __forceinline__ void fetch_point(const DevPlan& P, const PointSource& src, int64_t pt, float& x0, float& x1,
                                            float& x2) {
  if (src.x) {
    x0 = __ldg(src.x + pt * 3);
    x1 = __ldg(src.x + pt * 3 + 1);
    x2 = __ldg(src.x + pt * 3 + 2);
  } else {
    const int64_t ray = pt / src.S;
    const float zz = __ldg(src.z + pt);
    x0 = normalise1(P, 0, __fadd_rn(__ldg(src.rays_o + ray * 3 + 0), __fmul_rn(__ldg(src.rays_d + ray * 3 + 0), zz)));
    x1 = normalise1(P, 1, __fadd_rn(__ldg(src.rays_o + ray * 3 + 1), __fmul_rn(__ldg(src.rays_d + ray * 3 + 1), zz)));
    x2 = normalise1(P, 2, __fadd_rn(__ldg(src.rays_o + ray * 3 + 2), __fmul_rn(__ldg(src.rays_d + ray * 3 + 2), zz)));
  }
}

// losses from the (globally summed) statistics: JointEncodingNaruto.forward's train branch + get_sdf_loss
// (src/slam/coslam/model/scene_rep.py:244-287; tp/model/utils.py:103-148).  One thread.
__device__ __forceinline__ void finalize_losses(const double* stats, float* __restrict__ losses) {
  const double B = stats[NRT_STAT_N_RAYS], V = stats[NRT_STAT_N_VALID], NS = stats[NRT_STAT_N_SAMPLES];
  const double nfs = stats[NRT_STAT_N_FS], nsdf = stats[NRT_STAT_N_SDF];
  const double ntot = nfs + nsdf;
  // the reference evaluates these in fp32 tensors; fp64 here only removes summation-order noise
  float rgb_loss = (float)(stats[NRT_STAT_RGB_SQ] / (3.0 * B));
  float depth_loss = (float)(stats[NRT_STAT_DEPTH_SQ] / V);
  float fs_w = 1.0f - (float)nfs / (float)ntot;
  float sdf_w = 1.0f - (float)nsdf / (float)ntot;
  float fs_loss = (float)(stats[NRT_STAT_FS_SQ] / NS) * fs_w;
  float sdf_loss = (float)(stats[NRT_STAT_SDF_SQ] / NS) * sdf_w;
  float mean_inv2u = (float)(stats[NRT_STAT_INV2U] / V);
  float uncert_loss = mean_inv2u * depth_loss + 0.5f * (float)(stats[NRT_STAT_LOGU] / V);
  losses[NRT_LOSS_RGB] = rgb_loss;
  losses[NRT_LOSS_DEPTH] = depth_loss;
  losses[NRT_LOSS_SDF] = sdf_loss;
  losses[NRT_LOSS_FS] = fs_loss;
  losses[NRT_LOSS_UNCERT] = uncert_loss;
  losses[NRT_LOSS_PSNR] = -10.0f * logf(rgb_loss) / logf(10.0f);
  losses[NRT_LOSS_UNCERT_MIN] = (float)stats[NRT_STAT_UNCERT_MIN];
  losses[NRT_LOSS_RESERVED] = 0.f;
}

// error plumbing (api.cu)
void nrt_set_error(const char* fmt, ...);
// SM count of the current device (cached per device; 148 if there is none)
int nrt_device_sm_count();

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-DEVICE setting: one flag per (call site, device), so a process
// that drives several GPUs sets it on each of them.  Racing threads at worst set the attribute twice.
struct PerDeviceOnce {
  bool done[64] = {};
  bool first() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};
#define NRT_CUDA_CHECK(expr)                                                          \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      nrt_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return NRT_ERR_CUDA;                                                            \
    }                                                                                 \
  } while (0)
#define NRT_REQUIRE(cond, msg)                         \
  do {                                                 \
    if (!(cond)) {                                     \
      nrt_set_error("invalid argument: %s", msg);     \
      return NRT_ERR_INVALID;                          \
    }                                                  \
  } while (0)
