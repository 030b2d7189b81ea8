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

// ---------------------------------------------------------------------------------------------
// The same pass with NO look-up grids: everything the reference's constructor tabulates (C2E.grid, the six E2P.coor_xy, the
// back-projected texel rays) is a closed-form function of the output pixel, evaluated here per thread.  Per pixel (i, j):
//   longitude / latitude of the panorama pixel  ->  cube face by the reference's band / cap rule  ->  gnomonic face
//   coordinates  ->  nearest texel (ty, tx)  ->  that texel's tangent-plane point, rotated into the panorama frame by the
//   face's frame  ->  its longitude / latitude  ->  panorama coordinate  ->  4-tap bilinear read  ->  x |K^-1 [tx, ty, 1]|.
// Roundings follow the dtypes of the reference's constructor (numpy float32 angles widened to float64 coordinates, fp32
// torch ops for the tangent-plane grid) so that the nearest-texel choice agrees except where a coordinate sits within an
// ulp of a texel boundary (libm vs CUDA sinf / tanf / atan2f).  Drops 46 MB of grid traffic per panorama.
// ---------------------------------------------------------------------------------------------
struct ErpFrames {
  float rot[6][9];      // row-vector frames: p_pano = p_face @ rot[f]  (faces F R B L U D)
  float x_max;          // tan(fov / 2) of a 90-degree face in fp32
};

__device__ __forceinline__ double np_linspace(double start, double stop, int num, int i) {
  // numpy.linspace in float64: arange * step + start (two roundings), last element pinned to `stop`
  if (num > 1 && i == num - 1) return stop;
  const double step = num > 1 ? (stop - start) / (double)(num - 1) : 0.0;
  return __dadd_rn(__dmul_rn((double)i, step), start);
}

__global__ void __launch_bounds__(256) erp_depth2dist_analytic_kernel(const float* __restrict__ depth, int H, int W, int s,
                                                                      const ErpFrames F, float* __restrict__ dist) {
  const double PI = 3.141592653589793;
  const int64_t n = (int64_t)H * W;
  const int Wq = W / 4, shift = (3 * W) / 8;
  const float xm = F.x_max;
  const float lin_step = s > 1 ? __fdiv_rn(__fsub_rn(xm, -xm), (float)(s - 1)) : 0.f;
  for (int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pix < n; pix += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(pix / W), j = (int)(pix - (int64_t)i * W);
    const float lon = (float)np_linspace(-PI, PI, W, j);
    const float lat = __fmul_rn((float)np_linspace(PI, -PI, H, i), 0.5f);
    // ---- face: four longitude bands of W/4 columns starting at column 3W/8, with the ceiling / floor caps cut out along
    //      the rows where a band's edge latitude atan(cos(longitude / 4 ...)) falls (src/layers/c2e_utils.py:75-93) ----
    const int jj = ((j - shift) % W + W) % W;
    const int band = jj / Wq, col = jj - band * Wq;
    const double edge = np_linspace(-PI, PI, Wq, col) / 4.0;
    const int cap_rows = H / 2 - (int)rint(atan(cos(edge)) * (double)H / PI);
    const int face = (H - 1 - i) < cap_rows ? 5 : (i < cap_rows ? 4 : band);
    // ---- gnomonic coordinates on that face, fp32 like numpy's float32 arrays ----
    // The texel choice below is a rounding decision and ~10 % of the pixels of a small panorama sit within 1e-5 of a texel
    // boundary (symmetric lattices), so the fp32 tan / sin / cos are evaluated in double and rounded once: the correctly
    // rounded value, which is what the reference's libm-backed numpy float32 functions return (0 texel mismatches against
    // its grids on the golden cases, 6 of 2 M pixels at 1024 x 2048; CUDA's tanf / sinf, 2-4 ulp, flipped 0.6 %).
    float cx, cy;
    if (face < 4) {
      const float a = __fsub_rn(lon, (float)(PI * (double)face / 2.0));
      cx = __fmul_rn(0.5f, (float)tan((double)a));
      cy = __fdiv_rn(__fmul_rn(-0.5f, (float)tan((double)lat)), (float)cos((double)a));
    } else {
      const float c0 = __fmul_rn(0.5f, (float)tan((double)__fsub_rn((float)(PI / 2.0), face == 4 ? lat : fabsf(lat))));
      cx = __fmul_rn(c0, (float)sin((double)lon));
      cy = face == 4 ? __fmul_rn(c0, (float)cos((double)lon)) : __fmul_rn(-c0, (float)cos((double)lon));
    }
    // float64 from here to the normalised grid value, then fp32 (C2E.grid is `.float()` of a float64 tensor)
    const double sm1 = (double)(s - 1);
    const double X = __dmul_rn(__dadd_rn(fmin(fmax((double)cx, -0.5), 0.5), 0.5), sm1);
    const double Y = __dmul_rn(__dadd_rn(fmin(fmax((double)cy, -0.5), 0.5), 0.5), sm1);
    const float gx = (float)__dsub_rn(__dmul_rn(__ddiv_rn(X, sm1), 2.0), 1.0);
    const float gy = (float)__dsub_rn(__dmul_rn(__ddiv_rn(Y, sm1), 2.0), 1.0);
    // ---- C2E: nearest texel (ATen grid_sampler: unnormalise with align_corners, nearbyint); the clip above keeps it inside ----
    const int tx = min(max((int)nearbyintf(gs_unnorm(gx, s)), 0), s - 1);
    const int ty = min(max((int)nearbyintf(gs_unnorm(gy, s)), 0), s - 1);
    // ---- E2P: the texel's point on the face's tangent plane (torch.linspace ladders), rotated into the panorama frame ----
    const float px = linspace_at(-xm, xm, lin_step, s, tx);
    const float py = -linspace_at(-xm, xm, lin_step, s, ty);
    const float* R = F.rot[face];
    const float qx = __fmaf_rn(1.0f, R[6], __fmaf_rn(py, R[3], __fmul_rn(px, R[0])));
    const float qy = __fmaf_rn(1.0f, R[7], __fmaf_rn(py, R[4], __fmul_rn(px, R[1])));
    const float qz = __fmaf_rn(1.0f, R[8], __fmaf_rn(py, R[5], __fmul_rn(px, R[2])));
    const float plon = atan2f(qx, qz);
    const float plat = atan2f(qy, sqrtf(__fadd_rn(__fmul_rn(qx, qx), __fmul_rn(qz, qz))));
    // panorama coordinate (uv2coor), normalised for grid_sample, un-normalised again by it: the same fp32 op sequence
    const float ex = __fsub_rn(__fmul_rn(__fadd_rn(__fdiv_rn(plon, (float)(2.0 * PI)), 0.5f), (float)W), 0.5f);
    const float ey = __fsub_rn(__fmul_rn(__fadd_rn(__fdiv_rn(-plat, (float)PI), 0.5f), (float)H), 0.5f);
    const float nx = __fsub_rn(__fmul_rn(__fdiv_rn(ex, (float)(W - 1)), 2.0f), 1.0f);
    const float ny = __fsub_rn(__fmul_rn(__fdiv_rn(ey, (float)(H - 1)), 2.0f), 1.0f);
    const float x = fminf((float)(W - 1), fmaxf(gs_unnorm(nx, W), 0.f));
    const float y = fminf((float)(H - 1), fmaxf(gs_unnorm(ny, H), 0.f));
    const float x0f = floorf(x), y0f = floorf(y);
    const int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
    const float wx1 = x - x0f, wy1 = y - y0f;
    const float wx0 = (x0f + 1.0f) - x, wy0 = (y0f + 1.0f) - y;
    float d = 0.f;
    d = __fmaf_rn(__ldg(depth + (int64_t)y0 * W + x0), wx0 * wy0, d);
    if (x1 < W) d = __fmaf_rn(__ldg(depth + (int64_t)y0 * W + x1), wx1 * wy0, d);
    if (y1 < H) d = __fmaf_rn(__ldg(depth + (int64_t)y1 * W + x0), wx0 * wy1, d);
    if (x1 < W && y1 < H) d = __fmaf_rn(__ldg(depth + (int64_t)y1 * W + x1), wx1 * wy1, d);
    // ---- depth2dist: || depth * K^-1 [tx, ty, 1] ||, K = [[s/2, 0, s/2], [0, s/2, s/2], [0, 0, 1]] ----
    const float inv = __fdiv_rn(2.0f, (float)s);
    const float rx = __fmaf_rn(inv, (float)tx, -1.0f), ry = __fmaf_rn(inv, (float)ty, -1.0f);
    const float vx = __fmul_rn(d, rx), vy = __fmul_rn(d, ry), vz = d;
    dist[pix] = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz)));
  }
}

int launch_erp_depth2dist_analytic(const float* depth, int H, int W, int s, const float* face_rot, float x_max, float* dist,
                                   int sm_count, cudaStream_t st) {
  const int64_t n = (int64_t)H * W;
  if (n == 0) return NRT_OK;
  ErpFrames F;
  for (int f = 0; f < 6; ++f)
    for (int k = 0; k < 9; ++k) F.rot[f][k] = face_rot[f * 9 + k];
  F.x_max = x_max;
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count * 8;
  if (blocks > cap) blocks = cap;
  erp_depth2dist_analytic_kernel<<<(unsigned)blocks, 256, 0, st>>>(depth, H, W, s, F, dist);
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
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
