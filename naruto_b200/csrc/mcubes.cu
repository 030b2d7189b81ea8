// Iso-surface extraction: the reference's marching cubes
// (third_parties/coslam/external/NumpyMarchingCubes/marching_cubes/src/marching_cubes.cpp:418-462, called by save_mesh /
// save_uncert_mesh through src/slam/coslam/coslam_utils.py:145 on the dense SDF sweep), with the O(n^3) work on the device and
// a result that is identical to the reference's, vertex for vertex and index for index.
//
// What the reference computes: a DUAL grid -- every voxel centre (i,j,k) is a cell whose eight corners (i,j,k) +- 1/2 take
// the mean of the 2x2x2 voxels around them (trilerp at weight 1/2; a corner is invalid when any of the eight voxels is out of
// bounds, -inf or |d| >= truncation) -> cube case in Bourke's numbering -> jump / magnitude thresholds (10) -> linear
// interpolation on the cut edges -> a triangle soup in scan order (i, then j, then k) -> a sequential "first come" vertex
// merge on a 1e-5 hash grid (27-neighbourhood, scan order) -> degenerate and duplicate faces removed.
//
// Here:  corners_kernel   one thread per dual-grid corner: the 8-voxel sum in the reference's order (fp32, no FMA)
//        count_kernel     one thread per cell: validity, case, thresholds -> triangles of the cell (0..5)
//        scan             exclusive prefix sum over cells in scan order (three small kernels; deterministic)
//        emit_kernel      one thread per cell: interpolated vertices -> soup[first triangle of the cell ...]
//        host             the vertex merge and the two face filters are order-dependent by definition (the id of a vertex
//                         is the number of distinct vertices met before it); they run on the host over the compacted soup,
//                         O(#triangles), as in the reference.
// All arithmetic that decides a branch or produces a coordinate is written with explicit round-to-nearest intrinsics in the
// reference's operation order, so device and reference agree bit for bit.
#include <memory>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "common.cuh"

__device__ __constant__ unsigned long long kTriTable[256] = {
#include "mc_tables.inc"
};

// cube-case bit b <-> corner offset (marching_cubes.cpp:192-199); edge e joins corners kEdgeA[e], kEdgeB[e] (:234-245)
__device__ __constant__ int kCornerX[8] = {0, 1, 1, 0, 0, 1, 1, 0};
__device__ __constant__ int kCornerY[8] = {1, 1, 0, 0, 1, 1, 0, 0};
__device__ __constant__ int kCornerZ[8] = {0, 0, 0, 0, 1, 1, 1, 1};
__device__ __constant__ int kEdgeA[12] = {0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3};
__device__ __constant__ int kEdgeB[12] = {1, 2, 3, 0, 5, 6, 7, 4, 4, 5, 6, 7};

#define MC_INVALID __int_as_float(0x7fc00000)      // quiet NaN marks an invalid dual-grid corner

// corner (a,b,c), a in [0,nx], ... = mean of voxels (a-1..a, b-1..b, c-1..c), summed in trilerp()'s order (:108-115)
__global__ void __launch_bounds__(256) mc_corners_kernel(const float* __restrict__ vol, int nx, int ny, int nz, float truncation,
                                                         float* __restrict__ corner) {
  const int64_t n = (int64_t)(nx + 1) * (ny + 1) * (nz + 1);
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % (nz + 1)), b = (int)((t / (nz + 1)) % (ny + 1)), a = (int)(t / ((int64_t)(nz + 1) * (ny + 1)));
    float out = MC_INVALID;
    if (a >= 1 && a < nx && b >= 1 && b < ny && c >= 1 && c < nz) {
      const int ox[8] = {0, 1, 0, 0, 1, 0, 1, 1}, oy[8] = {0, 0, 1, 0, 1, 1, 0, 1}, oz[8] = {0, 0, 0, 1, 0, 1, 1, 1};
      float dist = 0.f;
      bool ok = true;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float d = __ldg(vol + ((int64_t)(a - 1 + ox[q]) * ny + (b - 1 + oy[q])) * nz + (c - 1 + oz[q]));
        ok = ok && (d != -INFINITY) && (fabsf(d) < truncation);          // NaN fails the comparison, like the reference
        dist = __fadd_rn(dist, __fmul_rn(0.125f, d));                    // (1/2 * 1/2 * 1/2) * d, exact; then the running sum
      }
      if (ok) out = dist;
    }
    corner[t] = out;
  }
}

struct McCell {
  float d[8];        // corner values, cube-case bit order
  unsigned cube;
  bool emit;
};

__device__ __forceinline__ McCell mc_cell(const float* __restrict__ corner, int ny, int nz, int i, int j, int k, float iso,
                                          float thresh) {
  McCell r;
  r.cube = 0u;
  r.emit = true;
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    r.d[b] = __ldg(corner + ((int64_t)(i + kCornerX[b]) * (ny + 1) + (j + kCornerY[b])) * (nz + 1) + (k + kCornerZ[b]));
    if (r.d[b] != r.d[b]) r.emit = false;                                 // an invalid corner: no triangles (:189)
    if (r.d[b] < iso) r.cube |= 1u << b;
  }
  if (!r.emit) return r;
  // jump and magnitude thresholds (:202-219); symmetric, so the pair order does not matter
#pragma unroll
  for (int a = 0; a < 8; ++a) {
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      if (__fmul_rn(r.d[a], r.d[b]) < 0.0f) {
        if (__fadd_rn(fabsf(r.d[a]), fabsf(r.d[b])) > thresh) r.emit = false;
      } else if (fabsf(__fsub_rn(r.d[a], r.d[b])) > thresh) {
        r.emit = false;
      }
    }
    if (fabsf(r.d[a]) > thresh) r.emit = false;
  }
  // edge mask: an edge is cut iff its two corners are on different sides; masks 0 and 255 emit nothing (:226)
  unsigned mask = 0u;
#pragma unroll
  for (int e = 0; e < 12; ++e)
    if (((r.cube >> kEdgeA[e]) ^ (r.cube >> kEdgeB[e])) & 1u) mask |= 1u << e;
  if (mask == 0u || mask == 255u) r.emit = false;
  return r;
}

__device__ __forceinline__ int mc_tri_count(unsigned cube) {
  const unsigned long long w = kTriTable[cube];
  int n = 0;
  while (n < 15 && ((w >> (4 * n)) & 0xFull) != 0xFull) ++n;
  return n / 3;
}

__global__ void __launch_bounds__(256) mc_count_kernel(const float* __restrict__ corner, int nx, int ny, int nz, float iso, float thresh,
                                                       int* __restrict__ count) {
  const int64_t n = (int64_t)nx * ny * nz;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(t % nz), j = (int)((t / nz) % ny), i = (int)(t / ((int64_t)nz * ny));
    const McCell c = mc_cell(corner, ny, nz, i, j, k, iso, thresh);
    count[t] = c.emit ? mc_tri_count(c.cube) : 0;
  }
}

// ---- exclusive scan over cells: per-block sums -> scan of the block sums (one block) -> per-element offsets ------------------
#define SCAN_ITEMS 2048        // elements per block (256 threads x 8)
__global__ void __launch_bounds__(256) mc_block_sum_kernel(const int* __restrict__ count, int64_t n, long long* __restrict__ bsum) {
  __shared__ long long s[8];
  const int64_t base = (int64_t)blockIdx.x * SCAN_ITEMS;
  long long v = 0;
  for (int q = 0; q < 8; ++q) {
    const int64_t t = base + threadIdx.x * 8 + q;
    if (t < n) v += count[t];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long tot = 0;
    for (int w = 0; w < 8; ++w) tot += s[w];
    bsum[blockIdx.x] = tot;
  }
}
__global__ void mc_scan_blocks_kernel(long long* __restrict__ bsum, int nb, long long* __restrict__ total) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  long long run = 0;
  for (int b = 0; b < nb; ++b) {
    const long long v = bsum[b];
    bsum[b] = run;
    run += v;
  }
  *total = run;
}

__device__ __forceinline__ void mc_vertex(float iso, const float* p1, const float* p2, float d1, float d2, float* out) {
  // vertexInterp (:122-141): three early-outs in this order, then the linear interpolation
  const float* pick = nullptr;
  if (fabsf(__fsub_rn(iso, d1)) < 0.00001f) pick = p1;
  else if (fabsf(__fsub_rn(iso, d2)) < 0.00001f) pick = p2;
  else if (fabsf(__fsub_rn(d1, d2)) < 0.00001f) pick = p1;
  if (pick) {
    out[0] = pick[0];
    out[1] = pick[1];
    out[2] = pick[2];
    return;
  }
  const float mu = __fdiv_rn(__fsub_rn(iso, d1), __fsub_rn(d2, d1));
#pragma unroll
  for (int a = 0; a < 3; ++a) out[a] = __fadd_rn(p1[a], __fmul_rn(mu, __fsub_rn(p2[a], p1[a])));
}

__global__ void __launch_bounds__(256) mc_emit_kernel(const float* __restrict__ corner, int nx, int ny, int nz, float iso, float thresh,
                                                      const int* __restrict__ count, const long long* __restrict__ bsum,
                                                      float* __restrict__ soup) {
  // one block handles SCAN_ITEMS consecutive cells; thread-level offsets by a block-wide scan of 8-element partial sums
  __shared__ long long s_off[256];
  const int64_t n = (int64_t)nx * ny * nz;
  const int64_t base = (int64_t)blockIdx.x * SCAN_ITEMS + threadIdx.x * 8;
  int c8[8];
  long long mine = 0;
  for (int q = 0; q < 8; ++q) {
    c8[q] = base + q < n ? count[base + q] : 0;
    mine += c8[q];
  }
  s_off[threadIdx.x] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long run = bsum[blockIdx.x];
    for (int t = 0; t < 256; ++t) {
      const long long v = s_off[t];
      s_off[t] = run;
      run += v;
    }
  }
  __syncthreads();
  long long off = s_off[threadIdx.x];
  for (int q = 0; q < 8; ++q) {
    if (c8[q] == 0) continue;
    const int64_t t = base + q;
    const int k = (int)(t % nz), j = (int)((t / nz) % ny), i = (int)(t / ((int64_t)nz * ny));
    const McCell c = mc_cell(corner, ny, nz, i, j, k, iso, thresh);
    const unsigned long long w = kTriTable[c.cube];
    for (int v = 0; v < 3 * c8[q]; ++v) {
      const int e = (int)((w >> (4 * v)) & 0xFull);
      const int a = kEdgeA[e], b = kEdgeB[e];
      // corner positions: voxel centre -+ 1/2 (:178-185)
      const float pa[3] = {(float)i + (kCornerX[a] ? 0.5f : -0.5f), (float)j + (kCornerY[a] ? 0.5f : -0.5f),
                           (float)k + (kCornerZ[a] ? 0.5f : -0.5f)};
      const float pb[3] = {(float)i + (kCornerX[b] ? 0.5f : -0.5f), (float)j + (kCornerY[b] ? 0.5f : -0.5f),
                           (float)k + (kCornerZ[b] ? 0.5f : -0.5f)};
      mc_vertex(iso, pa, pb, c.d[a], c.d[b], soup + (off * 3 + v) * 3);
    }
    off += c8[q];
  }
}

// ---------------------------------------------------------------------------------------------
// host side: the order-dependent clean-up of the soup (merge_close_vertices(approx) + remove_degenerate_faces +
// remove_duplicate_faces, marching_cubes.cpp:253-416), own implementation of the same procedure
// ---------------------------------------------------------------------------------------------
namespace {
struct Key3 {
  int x, y, z;
  bool operator==(const Key3& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct Key3Hash {
  size_t operator()(const Key3& k) const {
    return ((size_t)k.x * 73856093u) ^ ((size_t)k.y * 19349669u) ^ ((size_t)k.z * 83492791u);
  }
};
inline int sgnf(float v) { return (0.0f < v) - (v < 0.0f); }

struct McResult {
  std::vector<double> verts;            // 3 per vertex
  std::vector<unsigned long long> faces;   // 3 per face
};

void merge_soup(const std::vector<float>& soup, McResult& out) {
  const float cell = 0.00001f;
  const size_t n_v = soup.size() / 3;
  std::vector<unsigned> lookup(n_v);
  std::unordered_map<Key3, unsigned, Key3Hash> seen;
  seen.max_load_factor(0.6f);
  seen.reserve(n_v * 2);
  unsigned cnt = 0;
  for (size_t v = 0; v < n_v; ++v) {
    const float x = soup[3 * v], y = soup[3 * v + 1], z = soup[3 * v + 2];
    const Key3 c{(int)(x / cell + 0.5f * sgnf(x)), (int)(y / cell + 0.5f * sgnf(y)), (int)(z / cell + 0.5f * sgnf(z))};
    unsigned found = 0xffffffffu;
    for (int i = -1; i <= 1 && found == 0xffffffffu; ++i)
      for (int j = -1; j <= 1 && found == 0xffffffffu; ++j)
        for (int k = -1; k <= 1; ++k) {
          auto it = seen.find(Key3{c.x + i, c.y + j, c.z + k});
          if (it != seen.end()) {
            found = it->second;
            break;
          }
        }
    if (found == 0xffffffffu) {
      seen[c] = cnt;
      out.verts.push_back(x);
      out.verts.push_back(y);
      out.verts.push_back(z);
      lookup[v] = cnt++;
    } else {
      lookup[v] = found;
    }
  }
  // faces: drop degenerate ones, then duplicates (same vertex set; the first occurrence stays, unsorted)
  struct TriHash {
    size_t operator()(const std::vector<unsigned>& t) const {
      const size_t p[] = {73856093, 19349669, 83492791};
      size_t r = 0;
      for (unsigned i : t) r = r ^ (size_t)i * p[i % 3];
      return r;
    }
  };
  std::unordered_set<std::vector<unsigned>, TriHash> faces_seen;
  for (size_t f = 0; f + 2 < n_v; f += 3) {
    const unsigned a = lookup[f], b = lookup[f + 1], c = lookup[f + 2];
    if (a == b || a == c || b == c) continue;
    std::vector<unsigned> key{a, b, c};
    if (key[0] > key[1]) std::swap(key[0], key[1]);
    if (key[1] > key[2]) std::swap(key[1], key[2]);
    if (key[0] > key[1]) std::swap(key[0], key[1]);
    if (!faces_seen.insert(key).second) continue;
    out.faces.push_back(a);
    out.faces.push_back(b);
    out.faces.push_back(c);
  }
}
}  // namespace

int64_t mc_workspace_bytes(int nx, int ny, int nz) {
  const int64_t cells = (int64_t)nx * ny * nz, corners = (int64_t)(nx + 1) * (ny + 1) * (nz + 1);
  const int64_t nb = (cells + SCAN_ITEMS - 1) / SCAN_ITEMS;
  return corners * 4 + cells * 4 + (nb + 2) * 8 + 256;
}

// Runs the extraction; synchronises the stream (export path).  The handle owns the host-side result.
int mc_extract(const float* vol, int nx, int ny, int nz, float iso, float truncation, void* workspace, cudaStream_t st, void** handle) {
  const int64_t cells = (int64_t)nx * ny * nz, corners = (int64_t)(nx + 1) * (ny + 1) * (nz + 1);
  const int nb = (int)((cells + SCAN_ITEMS - 1) / SCAN_ITEMS);
  uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
  float* corner = reinterpret_cast<float*>(w);
  int* count = reinterpret_cast<int*>(w + corners * 4);
  long long* bsum = reinterpret_cast<long long*>(w + ((corners * 4 + cells * 4 + 255) / 256) * 256);
  long long* total_dev = bsum + nb;
  const float thresh = 10.0f;
  std::unique_ptr<McResult> res(new McResult());          // released to the caller only on success
  if (cells > 0) {
    const unsigned g1 = (unsigned)((corners + 255) / 256 < 148 * 16 ? (corners + 255) / 256 : 148 * 16);
    mc_corners_kernel<<<g1, 256, 0, st>>>(vol, nx, ny, nz, truncation, corner);
    const unsigned g2 = (unsigned)((cells + 255) / 256 < 148 * 16 ? (cells + 255) / 256 : 148 * 16);
    mc_count_kernel<<<g2, 256, 0, st>>>(corner, nx, ny, nz, iso, thresh, count);
    mc_block_sum_kernel<<<nb, 256, 0, st>>>(count, cells, bsum);
    mc_scan_blocks_kernel<<<1, 32, 0, st>>>(bsum, nb, total_dev);
    long long n_tri = 0;
    NRT_CUDA_CHECK(cudaMemcpyAsync(&n_tri, total_dev, sizeof(long long), cudaMemcpyDeviceToHost, st));
    NRT_CUDA_CHECK(cudaStreamSynchronize(st));
    if (n_tri > 0) {
      float* soup_dev = nullptr;
      NRT_CUDA_CHECK(cudaMalloc(&soup_dev, (size_t)n_tri * 9 * sizeof(float)));     // export path: sized by the result
      mc_emit_kernel<<<nb, 256, 0, st>>>(corner, nx, ny, nz, iso, thresh, count, bsum, soup_dev);
      std::vector<float> soup((size_t)n_tri * 9);
      cudaError_t e = cudaMemcpyAsync(soup.data(), soup_dev, soup.size() * sizeof(float), cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      cudaFree(soup_dev);
      if (e != cudaSuccess) {
        nrt_set_error("marching cubes: %s", cudaGetErrorString(e));
        return NRT_ERR_CUDA;
      }
      merge_soup(soup, *res);
    }
    NRT_CUDA_CHECK(cudaGetLastError());
  }
  *handle = res.release();
  return NRT_OK;
}

void mc_sizes(void* handle, int64_t* n_verts, int64_t* n_faces) {
  McResult* r = reinterpret_cast<McResult*>(handle);
  *n_verts = (int64_t)r->verts.size() / 3;
  *n_faces = (int64_t)r->faces.size() / 3;
}
void mc_copy(void* handle, double* verts, unsigned long long* faces) {
  McResult* r = reinterpret_cast<McResult*>(handle);
  if (verts) std::copy(r->verts.begin(), r->verts.end(), verts);
  if (faces) std::copy(r->faces.begin(), r->faces.end(), faces);
}
void mc_release(void* handle) { delete reinterpret_cast<McResult*>(handle); }
