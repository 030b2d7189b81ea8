// ERP depth -> ERP radial distance (reference: src/layers/erp_conversions.py:288-354 ERPDepth2Dist.forward, called once per
// simulator step at src/simulator/habitat_simulator.py:143 on the 1024 x 2048 panorama with a 512-texel skybox).
//
// The reference runs 6 x E2P (2-D bilinear grid_sample, border padding, align_corners) -> 6 x depth2dist (back-projection,
// torch.norm) -> C2E (3-D nearest grid_sample over the 6-face cube, zeros padding, align_corners): 6 x 512^2 texels are
// produced of which the nearest-neighbour resampling reads at most H x W.  Fused here into ONE pass over the output
// panorama: thread = output pixel -> nearest cube texel (face, y, x) -> that texel's ERP coordinate -> 4-tap bilinear read of
// the depth panorama -> scale by the texel's back-projected ray -> Euclidean norm.  The 9.4 MB of skybox intermediates never
// exist; traffic is the three static grids plus one read and one write of the panorama.  Index arithmetic follows ATen's
// grid_sampler (unnormalise with align_corners, clip for border padding, nearbyint for nearest) so texel choices are
// bit-identical; the interpolation itself is fp32 like the reference.
#include "common.cuh"

__device__ __forceinline__ float gs_unnorm(float c, int size) { return __fmul_rn(__fdiv_rn(__fadd_rn(c, 1.0f), 2.0f), (float)(size - 1)); }

__global__ void __launch_bounds__(256) erp_depth2dist_kernel(const float* __restrict__ depth, int H, int W,
                                                             const float* __restrict__ c2e_grid,   // [H,W,3]: x, y, face
                                                             const float* __restrict__ face_coor,  // [6,s,s,2]
                                                             const float* __restrict__ face_rays,  // [3,s*s]
                                                             int s, float* __restrict__ dist) {
  const int64_t n = (int64_t)H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    // ---- C2E: nearest texel of the cube volume [D=6, s, s] ----
    const float gx = __ldg(c2e_grid + i * 3), gy = __ldg(c2e_grid + i * 3 + 1), gz = __ldg(c2e_grid + i * 3 + 2);
    const float fx = nearbyintf(gs_unnorm(gx, s)), fy = nearbyintf(gs_unnorm(gy, s)), fz = nearbyintf(gs_unnorm(gz, 6));
    float out = 0.f;                                  // zeros padding
    if (fx >= 0.f && fx <= (float)(s - 1) && fy >= 0.f && fy <= (float)(s - 1) && fz >= 0.f && fz <= 5.f) {
      const int tx = (int)fx, ty = (int)fy, face = (int)fz;
      const int64_t texel = (int64_t)ty * s + tx;
      // ---- E2P: bilinear sample of the panorama at the texel's ERP coordinate, border padding ----
      const float cx = __ldg(face_coor + (((int64_t)face * s * s) + texel) * 2);
      const float cy = __ldg(face_coor + (((int64_t)face * s * s) + texel) * 2 + 1);
      float x = fminf((float)(W - 1), fmaxf(gs_unnorm(cx, W), 0.f));
      float y = fminf((float)(H - 1), fmaxf(gs_unnorm(cy, H), 0.f));
      const float x0f = floorf(x), y0f = floorf(y);
      const int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
      const float wx1 = x - x0f, wy1 = y - y0f;           // se - nw weights as ATen forms them: (ix_se - ix) etc.
      const float wx0 = (x0f + 1.0f) - x, wy0 = (y0f + 1.0f) - y;
      float d = 0.f;
      // ATen accumulates nw, ne, sw, se in this order, skipping taps outside the image
      d = __fmaf_rn(__ldg(depth + (int64_t)y0 * W + x0), wx0 * wy0, d);
      if (x1 < W) d = __fmaf_rn(__ldg(depth + (int64_t)y0 * W + x1), wx1 * wy0, d);
      if (y1 < H) d = __fmaf_rn(__ldg(depth + (int64_t)y1 * W + x0), wx0 * wy1, d);
      if (x1 < W && y1 < H) d = __fmaf_rn(__ldg(depth + (int64_t)y1 * W + x1), wx1 * wy1, d);
      // ---- depth2dist: || depth * (K^-1 [u, v, 1]) || ----
      const int64_t ss = (int64_t)s * s;
      const float px = __fmul_rn(d, __ldg(face_rays + texel)), py = __fmul_rn(d, __ldg(face_rays + ss + texel)),
                  pz = __fmul_rn(d, __ldg(face_rays + 2 * ss + texel));
      out = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz)));
    }
    dist[i] = out;
  }
}

int launch_erp_depth2dist(const float* depth, int H, int W, const float* c2e_grid, const float* face_coor, const float* face_rays,
                          int s, float* dist, int sm_count, cudaStream_t st) {
  const int64_t n = (int64_t)H * W;
  if (n == 0) return NRT_OK;
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count * 8;             // a whole number of waves of resident CTAs
  if (blocks > cap) blocks = cap;
  erp_depth2dist_kernel<<<(unsigned)blocks, 256, 0, st>>>(depth, H, W, c2e_grid, face_coor, face_rays, s, dist);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}
