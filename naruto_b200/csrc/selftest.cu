// Self-test of the tcgen05 building blocks in umma.cuh, run by tests/test_umma_selftest.py on the GPU:
// one CTA multiplies small fp32 matrices with exactly the operand layouts, descriptors, commit/wait and TMEM
// read-back the MLP kernels use, so a descriptor mistake shows up as a wrong product here, not as a wrong render.
//   mode 0: D[128][N] = A[128][K] * B[N][K]^T        (both operands read K-major; the forward/backward-data GEMMs)
//   mode 1: D[MA][N] = Y[128 pts][MA]^T * X[128 pts][N]   (transposed K-major operands written with conflict-free scalar
//           stores, row pitch = 1 mod 8; rows MA..127 of D are don't-care; the weight-gradient GEMMs; single pass)
//   mode 2: as mode 0 with A staged in tensor memory (tcgen05.st) instead of shared memory
// passes = 1: plain TF32;  passes = 3: the hi/lo split (A_hi*B_hi + A_lo*B_hi + A_hi*B_lo).
#include "common.cuh"
#include "umma.cuh"

using namespace umma;

__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(int mode, const float* __restrict__ a, const float* __restrict__ b,
                                                               int K, int N, int passes, float* __restrict__ d) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 8);
  const int t = threadIdx.x, warp = t >> 5;
  // operand A: mode 0 -> [K/4][128][4]; mode 1 -> [128/4][128][4].  operand B: mode 0 -> [K/4][N][4]; mode 1 -> [N/4][128][4]
  const int a_floats = mode != 1 ? K * 128 : 32 * ((K + 7) / 8 * 8 + 1) * 4 + 512;
  const int b_floats = mode != 1 ? K * N : 32 * (N + 1) * 4;
  float* a_hi = reinterpret_cast<float*>(smem_raw + 1024);
  float* a_lo = a_hi + a_floats;
  float* b_hi = a_lo + (passes == 3 ? a_floats : 0);
  float* b_lo = b_hi + b_floats;

  if (warp == 0) {
    if (t == 0) {
      mbar_init(bar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<256>(tmem_slot);
  }

  if (mode == 2) {
    // A goes to tensor memory: hi in columns [64, 64+K), lo in [64+K, 64+2K); B as in mode 0
    for (int i = t; i < N * (K / 4); i += 128) {
      const int n = i % N, c = i / N;
      const float4 v = *reinterpret_cast<const float4*>(b + n * K + 4 * c);
      if (passes == 3) sts4_split(b_hi + (c * N + n) * 4, b_lo + (c * N + n) * 4, v.x, v.y, v.z, v.w);
      else sts4(b_hi + (c * N + n) * 4, v.x, v.y, v.z, v.w);
    }
  } else if (mode == 0) {
    for (int c = 0; c < K / 4; ++c) {
      const float4 v = *reinterpret_cast<const float4*>(a + t * K + 4 * c);
      if (passes == 3) sts4_split(a_hi + (c * 128 + t) * 4, a_lo + (c * 128 + t) * 4, v.x, v.y, v.z, v.w);
      else sts4(a_hi + (c * 128 + t) * 4, v.x, v.y, v.z, v.w);
    }
    for (int i = t; i < N * (K / 4); i += 128) {
      const int n = i % N, c = i / N;
      const float4 v = *reinterpret_cast<const float4*>(b + n * K + 4 * c);
      if (passes == 3) sts4_split(b_hi + (c * N + n) * 4, b_lo + (c * N + n) * 4, v.x, v.y, v.z, v.w);
      else sts4(b_hi + (c * N + n) * 4, v.x, v.y, v.z, v.w);
    }
  } else {
    // weight-gradient form: thread p owns point p and scatters its values into transposed, K-major operands
    // yT[pts/4][RA][4], xT[pts/4][RB][4]; RA, RB = 1 (mod 8) make the scalar stores bank-conflict free
    const int MA = K, RA = (MA + 7) / 8 * 8 + 1, RB = N + 1;
    for (int j = 0; j < MA; ++j) a_hi[((t >> 2) * RA + j) * 4 + (t & 3)] = tf32_hi(a[t * MA + j]);
    for (int j = 0; j < N; ++j) b_hi[((t >> 2) * RB + j) * 4 + (t & 3)] = tf32_hi(b[t * N + j]);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_slot;
  if (mode == 2) {
    for (int c = 0; c < K / 8; ++c) {
      float hi[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float x = a[t * K + 8 * c + i];
        hi[i] = passes == 3 ? tf32_hi(x) : x;
        lo[i] = x - hi[i];
      }
      tmem_st8(tmem_addr(tbase, 32 * warp, 64 + 8 * c), hi);
      if (passes == 3) tmem_st8(tmem_addr(tbase, 32 * warp, 64 + K + 8 * c), lo);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  if (t == 0) {
    const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
    bool acc = false;
    for (int p = 0; p < passes; ++p) {
      const uint32_t as = p == 1 ? al : ah, bs = p == 2 ? bl : bh;
      if (mode == 2) {
        const uint32_t idesc = idesc_tf32(128, N, 0, 0);
        for (int ks = 0; ks < K / 8; ++ks) {
          mma_tf32_ts(tbase, tmem_addr(tbase, 0, 64 + (p == 1 ? K : 0) + 8 * ks), desc_kmajor(bs, N, 2 * ks), idesc, acc);
          acc = true;
        }
      } else if (mode == 0) {
        const uint32_t idesc = idesc_tf32(128, N, 0, 0);
        for (int ks = 0; ks < K / 8; ++ks) {
          mma_tf32_ss(tbase, desc_kmajor(as, 128, 2 * ks), desc_kmajor(bs, N, 2 * ks), idesc, acc);
          acc = true;
        }
      } else {
        const uint32_t idesc = idesc_tf32(128, N, 0, 0);
        const int RA = (K + 7) / 8 * 8 + 1, RB = N + 1;
        for (int ks = 0; ks < 128 / 8; ++ks) {
          mma_tf32_ss(tbase, desc_kmajor(as, RA, 2 * ks), desc_kmajor(bs, RB, 2 * ks), idesc, acc);
          acc = true;
        }
      }
    }
    mma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int cb = 0; cb < N; cb += 16) {
    float v[16];
    tmem_ld16(tmem_addr(tbase, 32 * warp, cb), v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) d[t * N + cb + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tbase);
}

// raw probe: operand images are copied verbatim into shared memory and multiplied with caller-chosen descriptor fields
__global__ void __launch_bounds__(128, 1) umma_raw_kernel(const float* __restrict__ a_img, int a_bytes, const float* __restrict__ b_img,
                                                          int b_bytes, int N, int ksteps, int a_mn, int b_mn, int a_lbo, int a_sbo,
                                                          int a_kstep, int b_lbo, int b_sbo, int b_kstep, float* __restrict__ d) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 8);
  const int t = threadIdx.x, warp = t >> 5;
  float* a_s = reinterpret_cast<float*>(smem_raw + 1024);
  float* b_s = a_s + a_bytes / 4;
  if (warp == 0) {
    if (t == 0) {
      mbar_init(bar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<256>(tmem_slot);
  }
  for (int i = t; i < a_bytes / 4; i += 128) a_s[i] = a_img[i];
  for (int i = t; i < b_bytes / 4; i += 128) b_s[i] = b_img[i];
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_slot;
  if (t == 0) {
    const uint32_t idesc = idesc_tf32(128, N, a_mn, b_mn);
    for (int ks = 0; ks < ksteps; ++ks)
      mma_tf32_ss(tbase, smem_desc(smem_u32(a_s) + ks * a_kstep, a_lbo, a_sbo), smem_desc(smem_u32(b_s) + ks * b_kstep, b_lbo, b_sbo),
                  idesc, ks > 0);
    mma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int cb = 0; cb < N; cb += 16) {
    float v[16];
    tmem_ld16(tmem_addr(tbase, 32 * warp, cb), v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) d[t * N + cb + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tbase);
}

int launch_umma_raw(const float* a_img, int a_bytes, const float* b_img, int b_bytes, int N, int ksteps, int a_mn, int b_mn,
                    int a_lbo, int a_sbo, int a_kstep, int b_lbo, int b_sbo, int b_kstep, float* d, cudaStream_t st) {
  NRT_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0 && ksteps >= 1, "raw probe N / ksteps");
  NRT_REQUIRE(a_bytes % 16 == 0 && b_bytes % 16 == 0 && a_bytes + b_bytes + 1024 <= 227 * 1024, "raw probe images");
  const size_t smem = 1024 + (size_t)a_bytes + b_bytes;
  NRT_CUDA_CHECK(cudaFuncSetAttribute(umma_raw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_raw_kernel<<<1, 128, smem, st>>>(a_img, a_bytes, b_img, b_bytes, N, ksteps, a_mn, b_mn, a_lbo, a_sbo, a_kstep, b_lbo, b_sbo,
                                        b_kstep, d);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}

int launch_umma_selftest(int mode, const float* a, const float* b, int K, int N, int passes, float* d, cudaStream_t st) {
  NRT_REQUIRE(mode >= 0 && mode <= 2, "selftest mode");
  NRT_REQUIRE(mode != 2 || (N <= 64 && K <= 96), "selftest mode 2: N <= 64, K <= 96");
  NRT_REQUIRE(passes == 1 || passes == 3, "selftest passes");
  NRT_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0, "selftest N");
  NRT_REQUIRE(K >= 8 && K % 8 == 0, "selftest K");
  NRT_REQUIRE(mode != 1 || (passes == 1 && K >= 8 && K <= 128 && K % 8 == 0), "selftest mode 1: single pass, 8 <= MA <= 128");
  const size_t a_floats = mode != 1 ? (size_t)K * 128 : 32 * ((K + 7) / 8 * 8 + 1) * 4 + 512,
               b_floats = mode != 1 ? (size_t)K * N : (size_t)32 * (N + 1) * 4;
  const size_t smem = 1024 + (a_floats + b_floats) * sizeof(float) * (passes == 3 ? 2 : 1);
  NRT_REQUIRE(smem <= 227 * 1024, "selftest operands exceed shared memory");
  NRT_CUDA_CHECK(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_selftest_kernel<<<1, 128, smem, st>>>(mode, a, b, K, N, passes, d);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}
