// extern "C" entry points of libnaruto_b200.so (see include/naruto_b200.h for the contract).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "common.cuh"

// launchers implemented in forward.cu / backward.cu / optim.cu
int launch_encode_fwd(const NrtPlan*, const float*, const float*, int64_t, float*, cudaStream_t);
int launch_oneblob_fwd(const float*, int64_t, float*, cudaStream_t);
int launch_oneblob_bwd(const float*, int64_t, const float*, float*, cudaStream_t);
int launch_decode_fwd(const NrtPlan*, const NrtParams*, const float*, int64_t, int, float*, float*, float*, cudaStream_t);
int launch_sample_z(const NrtPlan*, const float*, int64_t, const float*, int, uint64_t, float*, cudaStream_t);
int launch_render_fwd(const NrtPlan*, const NrtParams*, const float*, const float*, const float*, int64_t, const float*,
                      const float*, int, uint64_t, const NrtRenderOut*, const float*, double*, float*, const int*, cudaStream_t);
int launch_composite_fwd(const NrtPlan*, const float*, const float*, int64_t, int, const NrtRenderOut*, cudaStream_t);
int64_t loss_stats_doubles();
int launch_loss_partial(const NrtPlan*, const NrtRenderOut*, const float*, const float*, int64_t, double*, cudaStream_t);
int launch_loss_finalize(const double*, float*, cudaStream_t);
int launch_composite_bwd(const NrtPlan*, const NrtRenderOut*, const float*, const float*, int64_t, const double*, const float*,
                         float*, cudaStream_t);
int launch_decode_bwd(const NrtPlan*, const NrtParams*, const PointSource&, int64_t, const float*, int, const uint32_t*, const float*, float*,
                      const NrtGrads*, cudaStream_t);
int launch_decode_bwd_q(const NrtPlan*, const NrtParams*, const float*, const float*, const float*, int, int64_t, const float*,
                        const uint32_t*, const float*, const NrtGrads*, float*, cudaStream_t);
int64_t decode_bwd_q_scratch_floats(const NrtPlan*);
int launch_encode_bwd(const NrtPlan*, const float*, const PointSource&, int64_t, const float*, float, float*, float*,
                      cudaStream_t);
int q_trace_read(void*, int);
int ws_trace_read(void*, int);
int peer_trace_read(void*, int);
int launch_smooth(const NrtPlan*, const float*, const float*, int, double, double, float, float*, float*, void*, int, int, cudaStream_t);
int launch_adam(float*, float*, float*, float*, int64_t, int, const int*, float, float, float, float, float, int, int,
                cudaStream_t);
int launch_counter_add(int*, int, cudaStream_t);
int launch_step_begin(int*, int, uint64_t, float*, int*, cudaStream_t);
int launch_adam_groups(float*, float*, float*, float*, const NrtAdamGroup*, int, int, int, cudaStream_t);
int launch_map_volumes(const NrtPlan*, const NrtParams*, const int*, float*, float*, cudaStream_t);
int64_t mc_workspace_bytes(int, int, int);
int mc_extract(const float*, int, int, int, float, float, void*, cudaStream_t, void**);
void mc_sizes(void*, int64_t*, int64_t*);
void mc_copy(void*, double*, unsigned long long*);
void mc_release(void*);
int launch_goal_aggregate(const float*, const float*, const int*, const float*, int64_t, const float*, int, float, float, float,
                          float*, float*, int*, int, cudaStream_t);
int launch_erp_depth2dist(const float*, int, int, const float*, const float*, const float*, int, float*, int, cudaStream_t);
int launch_erp_depth2dist_analytic(const float*, int, int, int, const float*, float, float*, int, cudaStream_t);
int launch_stats_exchange(const NrtPeerTable*, double*, unsigned int*, float*, cudaStream_t);
int launch_adam_peers(const NrtPeerTable*, float*, float*, const NrtAdamGroup*, int, int64_t, float*, const unsigned int*, unsigned int*, int,
                      cudaStream_t);
int launch_camera_rays(int, int, float, float, float, float, float*, cudaStream_t);
int launch_pack_frame(const float*, const float*, const float*, int64_t, float*, cudaStream_t);
int launch_valid_depth_count(const float*, int64_t, float, int*, cudaStream_t);
int launch_kf_store(const float*, const int64_t*, int64_t, int, const int*, float*, cudaStream_t);
int launch_feistel_sample(int64_t, int64_t, uint64_t, const int*, int64_t*, cudaStream_t);
int launch_assemble_rays(const float*, const int64_t*, int, int, const int64_t*, int64_t, const float*, const int64_t*, int64_t,
                         const float*, int, float*, float*, float*, float*, cudaStream_t);
int launch_active_select(const float*, const float*, const float*, const float*, int64_t, int64_t, const float*, int, int, int,
                         const float*, int, int, int, float*, float*, float*, float*, int*, void*, cudaStream_t);
int launch_umma_selftest(int, const float*, const float*, int, int, int, float*, cudaStream_t);
int launch_umma_raw(const float*, int, const float*, int, int, int, int, int, int, int, int, int, int, int, float*, cudaStream_t);

static thread_local char g_err[512] = "";

int nrt_device_sm_count() {
  static int cached[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    (void)cudaGetLastError();
    return 148;
  }
  if (dev >= 0 && dev < 64 && cached[dev] > 0) return cached[dev];
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
    (void)cudaGetLastError();
    return 148;
  }
  if (dev >= 0 && dev < 64) cached[dev] = sms;
  return sms;
}

void nrt_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" {

const char* nrt_last_error(void) { return g_err; }
int nrt_abi_version(void) { return NRT_ABI_VERSION; }

int nrt_plan_create(const NrtConfig* cfg, NrtPlan** out) {
  NRT_REQUIRE(cfg && out, "null config/out");
  NRT_REQUIRE(cfg->abi_version == NRT_ABI_VERSION, "NrtConfig.abi_version mismatch");
  NRT_REQUIRE(cfg->n_levels == NRT_L && cfg->n_features == 2, "kernels are built for a 16-level x 2-feature hash grid");
  NRT_REQUIRE(cfg->n_bins == NRT_BINS, "kernels are built for OneBlob n_bins=16");
  NRT_REQUIRE(cfg->hidden_dim == NRT_H && cfg->hidden_dim_color == NRT_H && cfg->geo_feat_dim == NRT_GEO,
              "kernels are built for hidden_dim=32, hidden_dim_color=32, geo_feat_dim=15");
  NRT_REQUIRE(cfg->log2_hashmap_size >= 4 && cfg->log2_hashmap_size <= 28, "log2_hashmap_size out of range");
  NRT_REQUIRE(cfg->n_samples_d >= 0 && cfg->n_range_d >= 1, "n_samples_d >= 0 and n_range_d >= 1 required");
  NRT_REQUIRE(cfg->n_samples_d + cfg->n_range_d <= NRT_SMAX, "more than 256 samples per ray");
  NRT_REQUIRE(cfg->uncert_dims[0] > 0 && cfg->uncert_dims[1] > 0 && cfg->uncert_dims[2] > 0, "uncert_dims");
  NRT_REQUIRE(cfg->trunc > 0.f, "trunc must be positive");
  NrtPlan* p = new (std::nothrow) NrtPlan();
  NRT_REQUIRE(p != nullptr, "out of host memory");
  p->cfg = *cfg;
  DevPlan& d = p->dev;
  // level table: tcnn GridEncoding ctor (offset table) + grid_scale / grid_resolution, evaluated in fp64 and
  // rounded once to fp32 (the same rule as oracle/tcnn_shim.py so both sides hold identical bits)
  uint64_t offset = 0;
  for (int l = 0; l < NRT_L; ++l) {
    double sc = std::exp2((double)l * std::log2(cfg->per_level_scale)) * (double)cfg->base_resolution - 1.0;
    float scale = (float)sc;
    uint32_t res = (uint32_t)std::ceil(scale) + 1u;
    uint64_t dense = (uint64_t)res * res * res;
    const uint64_t cap = 0xFFFFFFFFull / 2;
    if (dense > cap) dense = cap;
    dense = (dense + 7) / 8 * 8;
    uint64_t size = dense < (1ull << cfg->log2_hashmap_size) ? dense : (1ull << cfg->log2_hashmap_size);
    d.lv[l].scale = scale;
    d.lv[l].res = res;
    d.lv[l].size = (uint32_t)size;
    d.lv[l].offset = (uint32_t)offset;
    d.lv[l].hashed = ((uint64_t)res * res * res > size) ? 1u : 0u;
    d.lv[l].res2 = res * res;
    d.lv[l].magic = (uint32_t)((1ull << 32) / size);
    // run merging in the backward scatter: window of 2^agg consecutive samples (cells coarse relative to the sample spacing)
    // (a window of 2 on the next levels, res <= 110, was measured neutral and is not used)
    d.lv[l].agg = res <= 32u ? 3u : res <= 64u ? 2u : 0u;
    {
      const uint64_t span = 1ull + res + (uint64_t)res * res;
      d.lv[l].lim = size > span ? (uint32_t)(size - span) : 0u;
      d.lv[l].pad_[0] = d.lv[l].pad_[1] = d.lv[l].pad_[2] = 0u;
    }
    offset += size;
  }
  if (offset >= (1ull << 31)) {
    delete p;
    NRT_REQUIRE(false, "hash table too large for 32-bit entry offsets");
  }
  p->n_grid_floats = (int64_t)offset * 2;
  for (int i = 0; i < 3; ++i) {
    d.bb_min[i] = cfg->bound_min[i];
    d.bb_ext[i] = cfg->bound_max[i] - cfg->bound_min[i];     // fp32 subtraction, as the tensor op
    d.ud[i] = cfg->uncert_dims[i];
  }
  d.trunc = cfg->trunc;
  d.sc_trunc = (float)((double)cfg->sc_factor * (double)cfg->trunc);
  d.near_z = cfg->near_z;
  d.far_z = cfg->far_z;
  d.depth_trunc = cfg->depth_trunc;
  d.range_d = cfg->range_d;
  d.n_d = cfg->n_samples_d;
  d.n_r = cfg->n_range_d;
  d.S = cfg->n_samples_d + cfg->n_range_d;
  d.step_u = d.n_d > 1 ? (d.far_z - d.near_z) / (float)(d.n_d - 1) : 0.f;
  d.step_r = d.n_r > 1 ? (d.range_d - (-d.range_d)) / (float)(d.n_r - 1) : 0.f;
  d.step_n = d.n_r > 1 ? (d.far_z - d.near_z) / (float)(d.n_r - 1) : 0.f;
  p->sm_count = nrt_device_sm_count();   // plan creation is legal without a device (CPU-side symbol/level tests): 148 then
  *out = p;
  return NRT_OK;
}

void nrt_plan_destroy(NrtPlan* plan) { delete plan; }

int nrt_plan_sizes(const NrtPlan* plan, int64_t* n_grid_floats, int32_t* n_samples, int32_t* n_enc_dims) {
  NRT_REQUIRE(plan, "null plan");
  if (n_grid_floats) *n_grid_floats = plan->n_grid_floats;
  if (n_samples) *n_samples = plan->dev.S;
  if (n_enc_dims) *n_enc_dims = NRT_ENC;
  return NRT_OK;
}

int nrt_plan_levels(const NrtPlan* plan, float* scale, int32_t* resolution, int32_t* size, int32_t* offset) {
  NRT_REQUIRE(plan, "null plan");
  for (int l = 0; l < NRT_L; ++l) {
    if (scale) scale[l] = plan->dev.lv[l].scale;
    if (resolution) resolution[l] = (int32_t)plan->dev.lv[l].res;
    if (size) size[l] = (int32_t)plan->dev.lv[l].size;
    if (offset) offset[l] = (int32_t)plan->dev.lv[l].offset;
  }
  return NRT_OK;
}

int nrt_encode_fwd(const NrtPlan* plan, const float* grid, const float* x, int64_t n, float* out, void* stream) {
  NRT_REQUIRE(plan && grid && (n == 0 || (x && out)) && n >= 0, "encode_fwd arguments");
  return launch_encode_fwd(plan, grid, x, n, out, (cudaStream_t)stream);
}

int nrt_encode_bwd(const NrtPlan* plan, const float* grid, const float* x, int64_t n, const float* dout, float* dgrid,
                   float* dx, void* stream) {
  NRT_REQUIRE(plan && grid && n >= 0 && (n == 0 || (x && dout)), "encode_bwd arguments");
  NRT_REQUIRE((reinterpret_cast<uintptr_t>(dgrid) & 15u) == 0, "dgrid must be 16-byte aligned (paired 16-byte reductions)");
  PointSource src{x, nullptr, nullptr, nullptr, 1};
  return launch_encode_bwd(plan, grid, src, n, dout, 1.0f, dgrid, dx, (cudaStream_t)stream);
}

int nrt_oneblob_fwd(const NrtPlan* plan, const float* x, int64_t n, float* out, void* stream) {
  NRT_REQUIRE(plan && n >= 0 && (n == 0 || (x && out)), "oneblob_fwd arguments");
  return launch_oneblob_fwd(x, n, out, (cudaStream_t)stream);
}

int nrt_oneblob_bwd(const NrtPlan* plan, const float* x, int64_t n, const float* dout, float* dx, void* stream) {
  NRT_REQUIRE(plan && n >= 0 && (n == 0 || (x && dout && dx)), "oneblob_bwd arguments");
  return launch_oneblob_bwd(x, n, dout, dx, (cudaStream_t)stream);
}

static int check_params(const NrtParams* p) {
  NRT_REQUIRE(p && p->grid && p->w1 && p->w2 && p->w3 && p->w4 && p->uncert, "NrtParams has a null tensor");
  return NRT_OK;
}

int nrt_decode_fwd(const NrtPlan* plan, const NrtParams* params, const float* x, int64_t n, int with_color, float* raw,
                   float* sdf_uncert, float* geo, void* stream) {
  NRT_REQUIRE(plan && n >= 0 && (n == 0 || x), "decode_fwd arguments");
  if (int rc = check_params(params)) return rc;
  NRT_REQUIRE(!(raw && !with_color), "raw output needs with_color=1");
  return launch_decode_fwd(plan, params, x, n, with_color, raw, sdf_uncert, geo, (cudaStream_t)stream);
}

int nrt_sample_z(const NrtPlan* plan, const float* target_d, int64_t n_rays, const float* u, int perturb, uint64_t seed,
                 float* z_vals, void* stream) {
  NRT_REQUIRE(plan && n_rays >= 0 && (n_rays == 0 || (target_d && z_vals)), "sample_z arguments");
  return launch_sample_z(plan, target_d, n_rays, u, perturb, seed, z_vals, (cudaStream_t)stream);
}

int nrt_render_fwd(const NrtPlan* plan, const NrtParams* params, const float* rays_o, const float* rays_d,
                   const float* target_d, int64_t n_rays, const float* z_in, const float* u, int perturb, uint64_t seed,
                   const NrtRenderOut* out, void* stream) {
  NRT_REQUIRE(plan && out && n_rays >= 0 && (n_rays == 0 || (rays_o && rays_d)), "render_fwd arguments");
  NRT_REQUIRE(z_in || target_d || n_rays == 0, "render_fwd needs target_d or z_in");
  if (int rc = check_params(params)) return rc;
  return launch_render_fwd(plan, params, rays_o, rays_d, target_d, n_rays, z_in, u, perturb, seed, out, nullptr, nullptr, nullptr, nullptr,
                           (cudaStream_t)stream);
}

int nrt_render_fwd_stats(const NrtPlan* plan, const NrtParams* params, const float* rays_o, const float* rays_d,
                         const float* target_rgb, const float* target_d, int64_t n_rays, const float* u, int perturb,
                         uint64_t seed, const int32_t* seed_step, const NrtRenderOut* out, double* stats, float* losses, void* stream) {
  NRT_REQUIRE(plan && out && rays_o && rays_d && target_rgb && target_d && stats && n_rays > 0, "render_fwd_stats arguments");
  NRT_REQUIRE(out->rgb && out->depth && out->uncert && out->z_vals && out->raw, "loss needs rgb, depth, uncert, z_vals, raw");
  if (int rc = check_params(params)) return rc;
  return launch_render_fwd(plan, params, rays_o, rays_d, target_d, n_rays, nullptr, u, perturb, seed, out, target_rgb, stats,
                           losses, seed_step, (cudaStream_t)stream);
}

int nrt_composite_fwd(const NrtPlan* plan, const float* raw, const float* z, int64_t n_rays, int32_t n_samples,
                      const NrtRenderOut* out, void* stream) {
  NRT_REQUIRE(plan && out && n_rays >= 0 && (n_rays == 0 || (raw && z)), "composite_fwd arguments");
  NRT_REQUIRE(n_samples >= 1 && n_samples <= NRT_SMAX, "composite_fwd: 1 <= n_samples <= 256");
  return launch_composite_fwd(plan, raw, z, n_rays, n_samples, out, (cudaStream_t)stream);
}

int64_t nrt_loss_stats_bytes(void) { return loss_stats_doubles() * (int64_t)sizeof(double); }

int nrt_loss_partial(const NrtPlan* plan, const NrtRenderOut* rend, const float* target_rgb, const float* target_d,
                     int64_t n_rays, double* stats, void* stream) {
  NRT_REQUIRE(plan && rend && target_rgb && target_d && stats && n_rays > 0, "loss_partial arguments");
  return launch_loss_partial(plan, rend, target_rgb, target_d, n_rays, stats, (cudaStream_t)stream);
}

int nrt_loss_finalize(const NrtPlan* plan, const double* stats, float* losses, void* stream) {
  NRT_REQUIRE(plan && stats && losses, "loss_finalize arguments");
  return launch_loss_finalize(stats, losses, (cudaStream_t)stream);
}

int nrt_loss_fwd(const NrtPlan* plan, const NrtRenderOut* rend, const float* target_rgb, const float* target_d,
                 int64_t n_rays, double* stats, float* losses, void* stream) {
  if (int rc = nrt_loss_partial(plan, rend, target_rgb, target_d, n_rays, stats, stream)) return rc;
  return nrt_loss_finalize(plan, stats, losses, stream);
}

int nrt_decode_bwd(const NrtPlan* plan, const NrtParams* params, const float* x, int64_t n, const float* draw,
                   const NrtGrads* grads, void* workspace, void* stream) {
  NRT_REQUIRE(plan && x && draw && grads && workspace && n > 0, "decode_bwd arguments");
  if (int rc = check_params(params)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  float* feat = reinterpret_cast<float*>(workspace);          // [n,32] recomputed hash features
  if (int rc = launch_encode_fwd(plan, params->grid, x, n, feat, st)) return rc;
  PointSource src{x, nullptr, nullptr, nullptr, 1};
  return launch_decode_bwd(plan, params, src, n, feat, 0, nullptr, draw, nullptr, grads, st);   // scatters into grads->grid itself
}

int64_t nrt_render_bwd_workspace(const NrtPlan* plan, int64_t n_rays) {
  if (!plan || n_rays < 0) return 0;
  // dL/d raw, padded to whole tiles of 128 points (the backward kernel fetches it tile-wise by TMA); feature gradients stay on chip
  // + one weight-gradient block per CTA, reduced in fixed order by the second launch of the backward
  return ((n_rays * plan->dev.S + 127) / 128 * 128 * 5 + decode_bwd_q_scratch_floats(plan)) * (int64_t)sizeof(float);
}

int nrt_render_bwd(const NrtPlan* plan, const NrtParams* params, const float* rays_o, const float* rays_d,
                   const float* target_rgb, const float* target_d, int64_t n_rays, const NrtRenderOut* rend,
                   const double* stats, const float* loss_grad, const NrtGrads* grads, void* workspace, void* stream) {
  NRT_REQUIRE(plan && rays_o && rays_d && target_rgb && target_d && rend && stats && loss_grad && grads && workspace && n_rays > 0,
              "render_bwd arguments");
  NRT_REQUIRE(rend->z_vals && rend->raw && rend->feat, "render_bwd needs z_vals, raw and feat saved by render_fwd");
  NRT_REQUIRE((reinterpret_cast<uintptr_t>(grads->grid) & 15u) == 0, "grads->grid must be 16-byte aligned (paired 16-byte reductions)");
  if (int rc = check_params(params)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_pts = n_rays * plan->dev.S;
  float* draw = reinterpret_cast<float*>(workspace);
  if (int rc = launch_composite_bwd(plan, rend, target_rgb, target_d, n_rays, stats, loss_grad, draw, st)) return rc;
  // saved masks: the 32-warp TMA-fed kernel (backward_q.cu); NRT_BWD_IMPL=tc selects round 1's kernel for A/B runs
  static const bool use_tc = [] {
    const char* e = getenv("NRT_BWD_IMPL");
    return e && strcmp(e, "tc") == 0;
  }();
  if (rend->masks && !use_tc)
    return launch_decode_bwd_q(plan, params, rays_o, rays_d, rend->z_vals, plan->dev.S, n_pts, rend->feat, rend->masks, draw, grads,
                               draw + (n_pts + 127) / 128 * 128 * 5, st);
  PointSource src{nullptr, rays_o, rays_d, rend->z_vals, plan->dev.S};
  return launch_decode_bwd(plan, params, src, n_pts, rend->feat, 1, rend->masks, draw, nullptr, grads, st);
}

int64_t nrt_smooth_workspace(const NrtPlan* plan, int32_t n) {
  if (!plan || n < 2) return 0;
  const int64_t m = n - 1;
  return m * m * m * NRT_ENC * (int64_t)sizeof(float);
}

int nrt_smooth_fwd_bwd(const NrtPlan* plan, const float* grid, const float* rand6, int32_t n, double voxel, double margin,
                       float loss_scale, float* loss, float* dgrid, void* workspace, int32_t part, int32_t n_parts, void* stream) {
  NRT_REQUIRE(plan && grid && rand6 && loss && workspace && n >= 2, "smooth arguments");
  NRT_REQUIRE(n_parts >= 1 && part >= 0 && part < n_parts, "smooth: 0 <= part < n_parts");
  NRT_REQUIRE((reinterpret_cast<uintptr_t>(dgrid) & 15u) == 0, "dgrid must be 16-byte aligned (paired 16-byte reductions)");
  return launch_smooth(plan, grid, rand6, n, voxel, margin, loss_scale, loss, dgrid, workspace, part, n_parts, (cudaStream_t)stream);
}

int nrt_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, int32_t step,
                  const int32_t* step_dev, float lr, float beta1, float beta2, float eps, float weight_decay, int zero_grad,
                  void* stream) {
  NRT_REQUIRE(param && grad && exp_avg && exp_avg_sq && n >= 0 && (step >= 1 || step_dev), "adam arguments");
  const int sms = nrt_device_sm_count();
  return launch_adam(param, grad, exp_avg, exp_avg_sq, n, step, step_dev, lr, beta1, beta2, eps, weight_decay, zero_grad, sms,
                     (cudaStream_t)stream);
}

int64_t nrt_mc_workspace_bytes(int32_t nx, int32_t ny, int32_t nz) {
  if (nx < 0 || ny < 0 || nz < 0) return 0;
  return mc_workspace_bytes(nx, ny, nz);
}

int nrt_mc_extract(const float* volume, int32_t nx, int32_t ny, int32_t nz, float isovalue, float truncation, void* workspace,
                   void* stream, void** result) {
  NRT_REQUIRE(result && nx >= 0 && ny >= 0 && nz >= 0, "mc_extract sizes");
  NRT_REQUIRE((int64_t)nx * ny * nz == 0 || (volume && workspace), "mc_extract arguments");
  return mc_extract(volume, nx, ny, nz, isovalue, truncation, workspace, (cudaStream_t)stream, result);
}

int nrt_mc_result_sizes(void* result, int64_t* n_vertices, int64_t* n_faces) {
  NRT_REQUIRE(result && n_vertices && n_faces, "mc_result_sizes arguments");
  mc_sizes(result, n_vertices, n_faces);
  return NRT_OK;
}

int nrt_mc_result_copy(void* result, double* vertices, uint64_t* faces) {
  NRT_REQUIRE(result, "mc_result_copy arguments");
  mc_copy(result, vertices, reinterpret_cast<unsigned long long*>(faces));
  return NRT_OK;
}

void nrt_mc_result_free(void* result) {
  if (result) mc_release(result);
}

int nrt_goal_aggregate(const float* uncert_vol, const float* sdf_vol, const int32_t* dims, const float* goal_pts, int64_t n_goal,
                       const float* topk_vxl, int32_t k, float min_dist, float max_dist, float safe_sdf, float* collections,
                       float* aggre, int32_t* n_valid, void* stream) {
  NRT_REQUIRE(dims && dims[0] > 0 && dims[1] > 0 && dims[2] > 0 && n_goal >= 0 && k >= 0, "goal_aggregate sizes");
  NRT_REQUIRE(n_goal == 0 || (uncert_vol && sdf_vol && goal_pts && aggre && (k == 0 || (topk_vxl && collections))),
              "goal_aggregate arguments");
  const int sms = nrt_device_sm_count();
  return launch_goal_aggregate(uncert_vol, sdf_vol, dims, goal_pts, n_goal, topk_vxl, k, min_dist, max_dist, safe_sdf, collections,
                               aggre, n_valid, sms, (cudaStream_t)stream);
}

int nrt_erp_depth2dist(const float* erp_depth, int32_t H, int32_t W, const float* c2e_grid, const float* face_coor,
                       const float* face_rays, int32_t skybox_size, float* erp_dist, void* stream) {
  NRT_REQUIRE(H >= 0 && W >= 0 && skybox_size >= 2, "erp_depth2dist sizes");
  NRT_REQUIRE((int64_t)H * W == 0 || (erp_depth && c2e_grid && face_coor && face_rays && erp_dist), "erp_depth2dist arguments");
  const int sms = nrt_device_sm_count();
  return launch_erp_depth2dist(erp_depth, H, W, c2e_grid, face_coor, face_rays, skybox_size, erp_dist, sms, (cudaStream_t)stream);
}

int nrt_erp_depth2dist_analytic(const float* erp_depth, int32_t H, int32_t W, int32_t skybox_size, const float* face_rot, float x_max,
                                float* erp_dist, void* stream) {
  NRT_REQUIRE(H >= 0 && W >= 0 && skybox_size >= 2 && face_rot, "erp_depth2dist_analytic sizes");
  NRT_REQUIRE((int64_t)H * W == 0 || (erp_depth && erp_dist && W % 4 == 0 && W >= 8 && H >= 2), "erp_depth2dist_analytic arguments");
  return launch_erp_depth2dist_analytic(erp_depth, H, W, skybox_size, face_rot, x_max, erp_dist, nrt_device_sm_count(),
                                        (cudaStream_t)stream);
}

static int check_peers(const NrtPeerTable* p) {
  NRT_REQUIRE(p && p->world >= 1 && p->world <= 8 && p->rank >= 0 && p->rank < p->world, "peer table: 1 <= world <= 8, 0 <= rank < world");
  for (int r = 0; r < p->world; ++r)
    NRT_REQUIRE(p->bucket[r] && p->theta[r] && p->stats_pad[r] && p->flags[r], "peer table has a null pointer");
  return NRT_OK;
}

int nrt_stats_exchange(const NrtPeerTable* peers, double* stats, uint32_t* xchg, float* losses, void* stream) {
  if (int rc = check_peers(peers)) return rc;
  NRT_REQUIRE(stats && xchg && losses, "stats_exchange arguments");
  return launch_stats_exchange(peers, stats, xchg, losses, (cudaStream_t)stream);
}

int nrt_adam_step_peers(const NrtPeerTable* peers, float* exp_avg, float* exp_avg_sq, const NrtAdamGroup* groups, int32_t n_groups,
                        int64_t smooth_slot, float* smooth_total, const uint32_t* xchg, uint32_t* done_counter, void* stream) {
  if (int rc = check_peers(peers)) return rc;
  NRT_REQUIRE(exp_avg && exp_avg_sq && groups && n_groups >= 1 && n_groups <= 3 && xchg && done_counter, "adam_step_peers arguments");
  for (int i = 0; i < n_groups; ++i)
    NRT_REQUIRE(groups[i].begin % 4 == 0 && groups[i].end >= groups[i].begin && (groups[i].step_dev || !groups[i].enabled),
                "adam group: begin a multiple of 4 floats, step counter on the device");
  return launch_adam_peers(peers, exp_avg, exp_avg_sq, groups, n_groups, smooth_slot, smooth_total, xchg, done_counter,
                           nrt_device_sm_count(), (cudaStream_t)stream);
}

int nrt_step_begin(int32_t* counter_dev, int32_t delta, uint64_t seed, float* rand6_dev, void* stream) {
  NRT_REQUIRE(counter_dev, "null counter");
  return launch_step_begin(counter_dev, delta, seed, rand6_dev, nullptr, (cudaStream_t)stream);
}

int nrt_iteration_begin(int32_t* map_counter_dev, int32_t* uncert_counter_dev, uint64_t seed, float* rand6_dev, void* stream) {
  NRT_REQUIRE(map_counter_dev, "null counter");
  return launch_step_begin(map_counter_dev, 1, seed, rand6_dev, uncert_counter_dev, (cudaStream_t)stream);
}

int nrt_adam_step_groups(float* param, float* grad, float* exp_avg, float* exp_avg_sq, const NrtAdamGroup* groups, int32_t n_groups,
                         int zero_grad, void* stream) {
  NRT_REQUIRE(param && grad && exp_avg && exp_avg_sq && groups && n_groups >= 1 && n_groups <= 3, "adam_step_groups arguments");
  NRT_REQUIRE(((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0,
              "adam_step_groups: the flat vectors must be 16-byte aligned");
  for (int i = 0; i < n_groups; ++i)
    NRT_REQUIRE(groups[i].begin % 4 == 0 && groups[i].end >= groups[i].begin && (groups[i].step_dev || !groups[i].enabled),
                "adam group: begin a multiple of 4 floats, step counter on the device");
  return launch_adam_groups(param, grad, exp_avg, exp_avg_sq, groups, n_groups, zero_grad, nrt_device_sm_count(), (cudaStream_t)stream);
}

int nrt_counter_add(int32_t* counter_dev, int32_t delta, void* stream) {
  NRT_REQUIRE(counter_dev, "null counter");
  return launch_counter_add(counter_dev, delta, (cudaStream_t)stream);
}

int nrt_map_volumes(const NrtPlan* plan, const NrtParams* params, const int32_t* dims, float* vol_uncert, float* vol_sdf, void* stream) {
  NRT_REQUIRE(plan && dims && vol_uncert && vol_sdf && dims[0] > 0 && dims[1] > 0 && dims[2] > 0, "map_volumes arguments");
  if (int rc = check_params(params)) return rc;
  return launch_map_volumes(plan, params, dims, vol_uncert, vol_sdf, (cudaStream_t)stream);
}

/* ---- device-resident ray sampling ---- */
int nrt_camera_rays(int32_t H, int32_t W, float fx, float fy, float cx, float cy, float* dirs, void* stream) {
  NRT_REQUIRE(H > 0 && W > 0 && dirs && fx != 0.f && fy != 0.f, "camera_rays arguments");
  return launch_camera_rays(H, W, fx, fy, cx, cy, dirs, (cudaStream_t)stream);
}

int nrt_pack_frame(const float* direction, const float* rgb, const float* depth, int64_t n_pixels, float* frame_rays, void* stream) {
  NRT_REQUIRE(n_pixels >= 0 && (n_pixels == 0 || (direction && rgb && depth && frame_rays)), "pack_frame arguments");
  return launch_pack_frame(direction, rgb, depth, n_pixels, frame_rays, (cudaStream_t)stream);
}

int nrt_valid_depth_count(const float* frame_rays, int64_t n_pixels, float depth_trunc, int32_t* count, void* stream) {
  NRT_REQUIRE(n_pixels >= 0 && count && (n_pixels == 0 || frame_rays), "valid_depth_count arguments");
  return launch_valid_depth_count(frame_rays, n_pixels, depth_trunc, count, (cudaStream_t)stream);
}

int nrt_kf_store(const float* frame_rays, const int64_t* idxs, int64_t n_idx, int32_t rays_per_kf, const int32_t* n_valid_dev, float* slot,
                 void* stream) {
  NRT_REQUIRE(frame_rays && idxs && slot && n_idx > 0 && rays_per_kf > 0, "kf_store arguments");
  return launch_kf_store(frame_rays, idxs, n_idx, rays_per_kf, n_valid_dev, slot, (cudaStream_t)stream);
}

int nrt_sample_indices(int64_t n, const int32_t* n_dev, int64_t k, uint64_t seed, int64_t* out, void* stream) {
  NRT_REQUIRE(k >= 0 && (k == 0 || out) && (n_dev || n > 0), "sample_indices arguments");
  return launch_feistel_sample(n, k, seed, n_dev, out, (cudaStream_t)stream);
}

int nrt_assemble_rays(const float* kf_rays, const int64_t* frame_ids, int32_t rays_per_kf, int32_t keyframe_every,
                      const int64_t* idxs_global, int64_t n_global, const float* cur_rays, const int64_t* idx_cur, int64_t n_cur,
                      const float* poses, int32_t n_poses, float* rays_o, float* rays_d, float* target_s, float* target_d,
                      void* stream) {
  NRT_REQUIRE(n_global >= 0 && n_cur >= 0 && rays_per_kf > 0 && keyframe_every > 0 && n_poses > 0 && poses, "assemble_rays arguments");
  NRT_REQUIRE(n_global == 0 || (kf_rays && frame_ids && idxs_global), "assemble_rays: global part");
  NRT_REQUIRE(n_cur == 0 || (cur_rays && idx_cur), "assemble_rays: current part");
  NRT_REQUIRE(n_global + n_cur == 0 || (rays_o && rays_d && target_s && target_d), "assemble_rays outputs");
  return launch_assemble_rays(kf_rays, frame_ids, rays_per_kf, keyframe_every, idxs_global, n_global, cur_rays, idx_cur, n_cur, poses,
                              n_poses, rays_o, rays_d, target_s, target_d, (cudaStream_t)stream);
}

int64_t nrt_active_select_workspace(int64_t n_rays) { return n_rays > 0 ? n_rays * (int64_t)sizeof(uint32_t) : 0; }

int nrt_active_select(const float* rays_o, const float* rays_d, const float* target_s, const float* target_d, int64_t n_rays,
                      int64_t n_cur, const float* uncert_vol, const int32_t* vol_dims, const float* bound_min, int32_t base_sample_num,
                      int32_t num_uncert_sample, int32_t oversample_mul, float* out_o, float* out_d, float* out_s, float* out_t,
                      int32_t* chosen, void* workspace, void* stream) {
  NRT_REQUIRE(rays_o && rays_d && target_s && target_d && uncert_vol && vol_dims && bound_min && out_o && out_d && out_s && out_t &&
                  workspace && n_rays > 0 && n_cur >= 0 && oversample_mul > 0,
              "active_select arguments");
  return launch_active_select(rays_o, rays_d, target_s, target_d, n_rays, n_cur, uncert_vol, vol_dims[0], vol_dims[1], vol_dims[2],
                              bound_min, base_sample_num, num_uncert_sample, oversample_mul, out_o, out_d, out_s, out_t, chosen,
                              workspace, (cudaStream_t)stream);
}

int nrt_debug_read(void* host_dst, int32_t bytes) {
  NRT_REQUIRE(host_dst && bytes != 0, "debug_read arguments");
  if (bytes > 0 && (bytes & (1 << 30))) return peer_trace_read(host_dst, bytes & ~(1 << 30));      // NRT_PEER_DEBUG=1 table
  if (bytes < 0) return ws_trace_read(host_dst, -bytes);     // negative size: the forward kernel's table (NRT_FWD_DEBUG=1)
  return q_trace_read(host_dst, bytes);
}

__global__ void stamp_kernel(unsigned long long* dst) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  *dst = t;
}

int nrt_debug_stamp(uint64_t* dst_dev, void* stream) {
  NRT_REQUIRE(dst_dev, "debug_stamp arguments");
  stamp_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned long long*>(dst_dev));
  NRT_CUDA_CHECK(cudaGetLastError());
  return NRT_OK;
}

int nrt_selftest_umma(int mode, const float* a, const float* b, int32_t k, int32_t n, int passes, float* d, void* stream) {
  NRT_REQUIRE(a && b && d, "selftest arguments");
  return launch_umma_selftest(mode, a, b, k, n, passes, d, (cudaStream_t)stream);
}

int nrt_selftest_umma_raw(const float* a_img, int32_t a_bytes, const float* b_img, int32_t b_bytes, int32_t n, int32_t ksteps,
                          int32_t a_mn, int32_t b_mn, int32_t a_lbo, int32_t a_sbo, int32_t a_kstep, int32_t b_lbo, int32_t b_sbo,
                          int32_t b_kstep, float* d, void* stream) {
  NRT_REQUIRE(a_img && b_img && d, "raw probe arguments");
  return launch_umma_raw(a_img, a_bytes, b_img, b_bytes, n, ksteps, a_mn, b_mn, a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep, d,
                         (cudaStream_t)stream);
}

}  // extern "C"
