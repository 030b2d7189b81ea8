/* naruto_b200 -- C-ABI of the B200-native mapping hot path (libnaruto_b200.so).
 *
 * Drop-in boundary for NARUTO's neural-implicit mapping iteration.  Every pointer marked "dev" is a
 * device pointer owned by the caller (in the reference's process: torch CUDA tensors, contiguous, fp32);
 * the library never allocates, frees or synchronises on the data path; every launch goes to the
 * cudaStream_t passed as `stream` (an opaque void* here so that the header needs no CUDA include).
 * All entry points return 0 on success or a negative NrtStatus; nrt_last_error() gives the text.
 *
 * Which reference interface each entry point replaces (paths relative to the NARUTO tree,
 * tp/ = third_parties/coslam/):
 *
 *   nrt_plan_create           tp/model/scene_rep.py:18-47 (get_resolution/get_encoding) +
 *                             tp/model/encodings.py:31-46,61-71 (tcnn.Encoding ctor arguments) +
 *                             src/slam/coslam/model/scene_rep.py:49-56 (get_uncert_grid dims)
 *   nrt_encode_fwd/_bwd       tcnn.Encoding HashGrid __call__ / autograd backward, reached from
 *                             src/slam/coslam/model/scene_rep.py:59,110 and tp/coslam.py:262 (smoothness)
 *   nrt_oneblob_fwd           tcnn.Encoding OneBlob __call__ (src/slam/coslam/model/scene_rep.py:118,144)
 *   nrt_decode_fwd            JointEncodingNaruto.query_sdf / query_color_sdf
 *                             (src/slam/coslam/model/scene_rep.py:98-148; decoder.py:29-41,99-116)
 *   nrt_sample_z              render_rays depth sampling (src/slam/coslam/model/scene_rep.py:158-180)
 *   nrt_render_fwd            JointEncodingNaruto.render_rays + raw2outputs + sdf2weights
 *                             (src/slam/coslam/model/scene_rep.py:150-225,66-96; tp/model/scene_rep.py:64-84)
 *   nrt_loss_fwd              JointEncodingNaruto.forward train branch + get_sdf_loss/get_masks
 *                             (src/slam/coslam/model/scene_rep.py:244-287; tp/model/utils.py:81-148)
 *   nrt_render_bwd            what autograd does for loss.backward() at src/slam/coslam/coslam.py:216,368
 *                             through the modules above
 *   nrt_smooth_fwd_bwd        CoSLAM.smoothness + its backward (tp/coslam.py:245-269)
 *   nrt_adam_step             torch.optim.Adam as configured at src/slam/coslam/coslam.py:409-419,240-243
 *   nrt_camera_rays .. nrt_active_select   the per-iteration ray sampling of global_BA (src/slam/coslam/coslam.py:302-359),
 *                             KeyFrameDatabase (tp/model/keyframe.py, src/slam/coslam/model/keyframe.py) and ActiveRaySampler
 *                             (src/slam/coslam/active_ray_sampler.py)
 */
#ifndef NARUTO_B200_H
#define NARUTO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NRT_ABI_VERSION 4

typedef enum NrtStatus {
  NRT_OK = 0,
  NRT_ERR_INVALID = -1,      /* bad argument / unsupported configuration */
  NRT_ERR_CUDA = -2,         /* a CUDA runtime call or kernel launch failed */
  NRT_ERR_NO_DEVICE = -3     /* no CUDA device / wrong architecture (needs sm_100) */
} NrtStatus;

/* The numbers JointEncodingNaruto reads out of the nested config dict, frozen at construction. */
typedef struct NrtConfig {
  int32_t abi_version;        /* = NRT_ABI_VERSION */
  /* tcnn HashGrid (tp/model/encodings.py:34-45) */
  int32_t n_levels;           /* 16 */
  int32_t n_features;         /* 2 */
  int32_t log2_hashmap_size;  /* config grid.hash_size */
  int32_t base_resolution;    /* 16 */
  double per_level_scale;     /* exp2(log2(resolution_sdf/16)/15) */
  /* tcnn OneBlob */
  int32_t n_bins;             /* 16 */
  /* decoder (src/slam/coslam/model/decoder.py) */
  int32_t hidden_dim;         /* 32 */
  int32_t geo_feat_dim;       /* 15 */
  int32_t hidden_dim_color;   /* 32 */
  /* scene */
  float bound_min[3];         /* bounding_box[:,0] as float32 */
  float bound_max[3];         /* bounding_box[:,1] as float32 */
  int32_t uncert_dims[3];     /* uncert_grid.shape = [Nx,Ny,Nz] */
  /* sampling / compositing */
  float trunc;                /* training.trunc */
  float sc_factor;            /* data.sc_factor */
  float near_z, far_z;        /* cam.near / cam.far */
  float depth_trunc;          /* cam.depth_trunc */
  int32_t n_samples_d;        /* training.n_samples_d */
  int32_t n_range_d;          /* training.n_range_d */
  float range_d;              /* training.range_d */
} NrtConfig;

typedef struct NrtPlan NrtPlan;   /* opaque */

/* Trainable tensors, in the reference's layouts (state_dict keys in comments). */
typedef struct NrtParams {
  const float* grid;     /* dev  embed_fn.params                    [n_entries*2] level-major, entry-major, feature-minor */
  const float* w1;       /* dev  decoder.sdf_net.model.0.weight     [32,80] row-major (out,in) */
  const float* w2;       /* dev  decoder.sdf_net.model.2.weight     [16,32] */
  const float* w3;       /* dev  decoder.color_net.model.0.weight   [32,63] */
  const float* w4;       /* dev  decoder.color_net.model.2.weight   [3,32]  */
  const float* uncert;   /* dev  uncert_grid                        [Nx,Ny,Nz] */
} NrtParams;

/* Gradient accumulators, same shapes; every backward entry point ADDS into them (torch .grad semantics). */
typedef struct NrtGrads {
  float* grid;
  float* w1;
  float* w2;
  float* w3;
  float* w4;
  float* uncert;
} NrtGrads;

/* Per-ray outputs of render_rays.  Any pointer may be NULL (not materialised). */
typedef struct NrtRenderOut {
  float* rgb;         /* dev [B,3] */
  float* depth;       /* dev [B]   */
  float* depth_var;   /* dev [B]   */
  float* acc;         /* dev [B]   acc_map   */
  float* disp;        /* dev [B]   disp_map  */
  float* uncert;      /* dev [B]   uncert_map */
  float* z_vals;      /* dev [B,S] */
  float* raw;         /* dev [B,S,5] = rgb logits(3), sdf, raw uncertainty */
  float* weights;     /* dev [B,S] */
  float* feat;        /* dev, ceil(B*S/128)*128*32 floats: hash features saved for nrt_render_bwd (training), opaque tile-major
                       * layout [tile of 128 points][8 chunks][128 points][4 floats] (coalesced for writer and reader) */
  uint32_t* masks;    /* dev [ceil(B*S/128)*128, 2] (padded to whole tiles like feat; rows >= B*S are never read as data) ReLU
                       * masks of the two hidden layers (bit j = unit j active), saved for nrt_render_bwd: with them the
                       * backward recomputes activations in single-pass TF32 (they only feed the tf32 weight-gradient
                       * operands) instead of 3xTF32 and fetches its tiles by TMA; optional (NULL: masks are recomputed exactly) */
} NrtRenderOut;

#define NRT_N_LOSS 8
/* losses[] layout written by nrt_loss_fwd: */
enum { NRT_LOSS_RGB = 0, NRT_LOSS_DEPTH = 1, NRT_LOSS_SDF = 2, NRT_LOSS_FS = 3, NRT_LOSS_UNCERT = 4,
       NRT_LOSS_PSNR = 5, NRT_LOSS_UNCERT_MIN = 6 /* min over rays of uncert_map (the reference asserts > 0) */,
       NRT_LOSS_RESERVED = 7 };

#define NRT_N_STATS 16
/* stats[] (device, fp64) layout shared by nrt_loss_fwd / nrt_render_bwd; a multi-GPU caller sums the
 * first NRT_N_STATS_SUM entries across ranks (nrt_loss_partial -> all-reduce -> nrt_loss_finalize). */
enum { NRT_STAT_N_RAYS = 0, NRT_STAT_N_VALID = 1, NRT_STAT_N_FS = 2, NRT_STAT_N_SDF = 3, NRT_STAT_N_SAMPLES = 4,
       NRT_STAT_RGB_SQ = 5, NRT_STAT_DEPTH_SQ = 6, NRT_STAT_FS_SQ = 7, NRT_STAT_SDF_SQ = 8,
       NRT_STAT_INV2U = 9, NRT_STAT_LOGU = 10, NRT_N_STATS_SUM = 11, NRT_STAT_UNCERT_MIN = 11 };

const char* nrt_last_error(void);
int nrt_abi_version(void);

/* ---- plan ------------------------------------------------------------------------------------ */
int nrt_plan_create(const NrtConfig* cfg, NrtPlan** out);
void nrt_plan_destroy(NrtPlan* plan);
/* sizes derived from the config: n_grid_floats = sum(level sizes)*n_features, n_samples = n_samples_d+n_range_d */
int nrt_plan_sizes(const NrtPlan* plan, int64_t* n_grid_floats, int32_t* n_samples, int32_t* n_enc_dims);
/* host arrays of length n_levels: the level table (tcnn grid_scale/grid_resolution/offset table) */
int nrt_plan_levels(const NrtPlan* plan, float* scale, int32_t* resolution, int32_t* size, int32_t* offset);

/* ---- encodings (lower seam: tcnn.Encoding) --------------------------------------------------- */
/* x: dev [n,3] already normalised to the bound; out: dev [n,32] */
int nrt_encode_fwd(const NrtPlan* plan, const float* grid, const float* x, int64_t n, float* out, void* stream);
/* dgrid += scatter(dout); dx (dev [n,3], may be NULL) = d out / d x contracted with dout (overwritten) */
int nrt_encode_bwd(const NrtPlan* plan, const float* grid, const float* x, int64_t n, const float* dout,
                   float* dgrid, float* dx, void* stream);
/* out: dev [n,48] */
int nrt_oneblob_fwd(const NrtPlan* plan, const float* x, int64_t n, float* out, void* stream);
/* dx: dev [n,3] (overwritten) */
int nrt_oneblob_bwd(const NrtPlan* plan, const float* x, int64_t n, const float* dout, float* dx, void* stream);

/* ---- point decode (query_sdf / query_color_sdf) ---------------------------------------------- */
/* x: dev [n,3] normalised.  raw: dev [n,5] or NULL; sdf_uncert: dev [n,2] or NULL; geo: dev [n,15] or NULL.
 * with_color = 0 skips the colour MLP (query_sdf); raw then requires with_color = 1. */
int nrt_decode_fwd(const NrtPlan* plan, const NrtParams* params, const float* x, int64_t n, int with_color,
                   float* raw, float* sdf_uncert, float* geo, void* stream);

/* get_map_volumes (src/slam/coslam/coslam_utils.py:58-97): the dense sweep the planner consumes.  dims: HOST int32[3] =
 * lattice points per axis (getVoxels: round(extent/voxel + 0.0005) + 1); the lattice is torch.linspace over the plan's
 * bound per axis, meshgrid 'ij'.  vol_sdf dev [d0,d1,d2]; vol_uncert dev [d0,d1,d2] = softplus(uncert)+0.01 where
 * 0 <= sdf < 0.5, else 0.  One launch: lattice generation, encode, SDF net and masking fused; the reference's discarded
 * `embed` pass is not reproduced. */
int nrt_map_volumes(const NrtPlan* plan, const NrtParams* params, const int32_t* dims, float* vol_uncert, float* vol_sdf,
                    void* stream);

/* ---- rays ------------------------------------------------------------------------------------ */
/* z_vals: dev [B,S].  u: dev [B,S] uniform draws (the reference's torch.rand) or NULL;
 * when u == NULL and perturb != 0 the kernel draws its own Philox stream from `seed`. */
int nrt_sample_z(const NrtPlan* plan, const float* target_d, int64_t n_rays, const float* u, int perturb,
                 uint64_t seed, float* z_vals, void* stream);

/* rays_o, rays_d: dev [B,3]; target_d: dev [B] (may be NULL iff z_in != NULL).
 * z_in: dev [B,S] externally sampled depths (parity mode) or NULL (sampled in-kernel as nrt_sample_z). */
int nrt_render_fwd(const NrtPlan* plan, const NrtParams* params, const float* rays_o, const float* rays_d,
                   const float* target_d, int64_t n_rays, const float* z_in, const float* u, int perturb,
                   uint64_t seed, const NrtRenderOut* out, void* stream);

/* nrt_render_fwd followed by nrt_loss_partial in ONE launch (the training path: JointEncodingNaruto.forward,
 * src/slam/coslam/model/scene_rep.py:227-287 = render_rays + the loss sums of tp/model/utils.py:81-148): the
 * compositing warps accumulate the shard's loss statistics while the ray's samples are still on chip.  `stats` is the
 * buffer of nrt_loss_partial (nrt_loss_stats_bytes() bytes, zero-filled once at allocation) and receives the same
 * NRT_N_STATS sums; out must provide rgb, depth, uncert, z_vals, raw (and feat for nrt_render_bwd).  With u == NULL the
 * stratified jitter (torch.rand(z_vals.shape), src/slam/coslam/model/scene_rep.py:176-180) is drawn in the kernel from
 * Philox keyed by seed; seed_step (optional, dev int32) is mixed into the key at launch, so a replayed CUDA graph draws new
 * jitter every iteration (pass the step counter advanced by nrt_step_begin).  losses (optional, dev fp32 [NRT_N_LOSS]):
 * when the shard is the whole batch (one GPU) the last CTA also writes what nrt_loss_finalize would; a multi-GPU caller
 * passes NULL, sums stats across ranks and calls nrt_loss_finalize. */
int nrt_render_fwd_stats(const NrtPlan* plan, const NrtParams* params, const float* rays_o, const float* rays_d,
                         const float* target_rgb, const float* target_d, int64_t n_rays, const float* u, int perturb,
                         uint64_t seed, const int32_t* seed_step, const NrtRenderOut* out, double* stats, float* losses,
                         void* stream);

/* raw2outputs + sdf2weights on caller-provided samples (JointEncodingNaruto.raw2outputs,
 * src/slam/coslam/model/scene_rep.py:66-96): raw dev [B,n_samples,5], z dev [B,n_samples]; fills the per-ray
 * fields and `weights` of out. */
int nrt_composite_fwd(const NrtPlan* plan, const float* raw, const float* z, int64_t n_rays, int32_t n_samples,
                      const NrtRenderOut* out, void* stream);

/* Loss statistics of one ray shard.  stats: dev fp64 buffer of nrt_loss_stats_bytes() bytes, zero-filled once
 * by the caller at allocation; its first NRT_N_STATS doubles are overwritten with this shard's sums (the rest
 * is scratch for the deterministic cross-block reduction). */
int64_t nrt_loss_stats_bytes(void);
int nrt_loss_partial(const NrtPlan* plan, const NrtRenderOut* rend, const float* target_rgb, const float* target_d,
                     int64_t n_rays, double* stats, void* stream);
/* losses (dev fp32 [NRT_N_LOSS]) from (globally summed) stats. */
int nrt_loss_finalize(const NrtPlan* plan, const double* stats, float* losses, void* stream);
/* convenience: partial + finalize for the single-GPU case */
int nrt_loss_fwd(const NrtPlan* plan, const NrtRenderOut* rend, const float* target_rgb, const float* target_d,
                 int64_t n_rays, double* stats, float* losses, void* stream);

/* Backward of the point decode alone (autograd of query_color_sdf / run_network): draw: dev [n,5] = dL/d raw.
 * Adds into grads; workspace: dev scratch of n*32*4 bytes. */
int nrt_decode_bwd(const NrtPlan* plan, const NrtParams* params, const float* x, int64_t n, const float* draw,
                   const NrtGrads* grads, void* workspace, void* stream);

/* Backward of total = sum_k loss_grad[k] * losses[k] (k < 5; loss_grad is a dev fp32 [5] array holding
 * dL/d{rgb,depth,sdf,fs,uncert}_loss, i.e. the reference's loss weights when called from get_loss_from_ret).
 * rend must hold z_vals, raw, feat and the per-ray rgb/depth/uncert written by nrt_render_fwd.
 * workspace: dev scratch of nrt_render_bwd_workspace(n_rays) bytes. */
int64_t nrt_render_bwd_workspace(const NrtPlan* plan, int64_t n_rays);
int nrt_render_bwd(const NrtPlan* plan, const NrtParams* params, const float* rays_o, const float* rays_d,
                   const float* target_rgb, const float* target_d, int64_t n_rays, const NrtRenderOut* rend,
                   const double* stats, const float* loss_grad, const NrtGrads* grads, void* workspace,
                   void* stream);

/* ---- smoothness ------------------------------------------------------------------------------ */
/* TV loss of the hash features on the (n-1)^3 lattice of CoSLAM.smoothness (n = smooth_pts, pitch `voxel`,
 * border `margin`).  rand6: dev fp32 [6] = the reference's two uniform draws, torch.rand(3) (offset) then
 * torch.rand((1,1,1,3)) (jitter); kept on the device so a captured CUDA graph can be replayed with fresh draws.
 * loss (dev fp32 [1], overwritten) = tv / n^3;  dgrid += loss_scale * d loss / d grid (dgrid may be NULL).
 * workspace: dev scratch of nrt_smooth_workspace(n) bytes.
 * part / n_parts: the term is independent of the rays, so data-parallel ranks split it: every rank encodes the whole
 * lattice but forms the loss and scatters the gradient only for slab `part` of `n_parts` of the lattice's scan order; the
 * slabs add up to the whole term (sum loss and dgrid over ranks, e.g. inside the gradient all-reduce).  0 / 1 = all of it. */
int64_t nrt_smooth_workspace(const NrtPlan* plan, int32_t n);
int nrt_smooth_fwd_bwd(const NrtPlan* plan, const float* grid, const float* rand6, int32_t n, double voxel,
                       double margin, float loss_scale, float* loss, float* dgrid, void* workspace, int32_t part,
                       int32_t n_parts, void* stream);

/* ---- optimiser ------------------------------------------------------------------------------- */
/* torch.optim.Adam (no amsgrad): grad += weight_decay * p; m, v update; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps).
 * step = 1-based step count of this update; if step_dev != NULL the count is read from that device int32
 * instead (graph replay; advance it with nrt_counter_add).  zero_grad != 0 clears the gradient in the same pass. */
int nrt_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, int32_t step,
                  const int32_t* step_dev, float lr, float beta1, float beta2, float eps, float weight_decay,
                  int zero_grad, void* stream);
int nrt_counter_add(int32_t* counter_dev, int32_t delta, void* stream);
/* nrt_counter_add plus the six uniforms of the smoothness lattice (torch.rand(3), torch.rand((1,1,1,3)),
 * tp/coslam.py:252-258) from Philox keyed by (seed, new counter value) -> rand6_dev (dev fp32 [6], optional). */
int nrt_step_begin(int32_t* counter_dev, int32_t delta, uint64_t seed, float* rand6_dev, void* stream);

/* ---- data-parallel exchanges over NVLink peer memory (no NCCL on the iteration's path) ----------------------------
 * Ray-sharded data parallelism exchanges the loss statistics before the backward pass and the gradient bucket
 * before Adam.  With these two entry points both exchanges happen inside the kernels that consume the data, over buffers that
 * every rank has mapped from every other rank (symmetric memory; the host side sets this up once, e.g. with
 * torch.distributed._symmetric_memory).  All pointer tables are HOST arrays of DEVICE pointers indexed by rank.
 *   bucket[r]     dev fp32 [pad4(total) + 4]  flat gradient [grid | w1 | w2 | w3 | w4 | uncert], the smoothness-loss slot behind it
 *   theta[r]      dev fp32 [pad4(total)]      flat parameters, same layout
 *   stats_pad[r]  dev u64  [world * 2 * NRT_N_STATS], zero-filled once (flag-carrying words: 32 data bits + exchange number)
 *   flags[r]      dev u32  [3 * 8], zero-filled once
 * Every rank must make the same sequence of calls (one nrt_stats_exchange, then one nrt_adam_step_peers per iteration). */
typedef struct NrtPeerTable {
  int32_t world, rank;        /* world <= 8 */
  float* bucket[8];
  float* theta[8];
  double* stats_pad[8];
  uint32_t* flags[8];
  float* bucket_mc;           /* optional: NVSwitch multicast (NVLS) address of the same bucket region on all ranks, or NULL; with it */
  float* theta_mc;            /* the gradient sum is one in-switch multimem.ld_reduce and the parameter all-gather one multimem.st */
} NrtPeerTable;

/* One Adam parameter group of the flat vector: floats [begin, end), begin a multiple of 4 and every buffer padded to a whole
 * number of float4s behind `end` (padding gradients stay zero), torch.optim.Adam hyper-parameters, the
 * group's device-side 1-based step counter (already advanced), enabled = 0 skips the group (its gradient keeps accumulating
 * locally, like the every-5th uncertainty-grid step of src/slam/coslam/coslam.py:397-399). */
typedef struct NrtAdamGroup {
  int64_t begin, end;
  float lr, beta1, beta2, eps, weight_decay;
  const int32_t* step_dev;
  int32_t enabled;
  int32_t keep_grad;   /* nrt_adam_step_peers only: 1 = do not clear the gradients that were read on the peers (the owner of each
                        * bucket clears its own copy before the next accumulation: no zero-stores over NVLink) */
} NrtAdamGroup;

/* all-reduce of the loss statistics + nrt_loss_finalize in one launch: stats (dev, this rank's sums from
 * nrt_render_fwd_stats) is overwritten with the global sums, losses as nrt_loss_finalize.  The statistics cross NVLink as
 * 8-byte words that carry the exchange number next to 32 data bits (no fence, no separate flag).  xchg: dev u32 exchange counter of
 * this rank (zero-filled once; advanced here, read by nrt_adam_step_peers). */
int nrt_stats_exchange(const NrtPeerTable* peers, double* stats, uint32_t* xchg, float* losses, void* stream);
/* reduce-scatter + Adam + all-gather in one launch: every rank reduces and steps its 1/world slice of each enabled group
 * (gradients summed over the ranks in rank order -- or inside the NVSwitch when peers->bucket_mc is set -- then cleared on every
 * rank unless the group's keep_grad is set), with its local moments exp_avg / exp_avg_sq
 * (dev fp32 [total]; only the rank's own slices are maintained), and stores the updated parameters into every rank's theta.
 * smooth_slot: index of the smoothness-loss slot in bucket (its sum over ranks -> smooth_total, dev fp32 [1], may be NULL).
 * done_counter: dev u32 scratch, zero-filled once.  Returns after every rank's stores are visible everywhere. */
int nrt_adam_step_peers(const NrtPeerTable* peers, float* exp_avg, float* exp_avg_sq, const NrtAdamGroup* groups, int32_t n_groups,
                        int64_t smooth_slot, float* smooth_total, const uint32_t* xchg, uint32_t* done_counter, void* stream);

/* The same three groups on ONE rank in one launch (torch.optim.Adam.step() + zero_grad() of create_optimizer's two groups,
 * src/slam/coslam/coslam.py:409-419, and of init_uncert_grid_optim's, :240-243): param / grad / exp_avg / exp_avg_sq are the
 * flat vectors [grid | w1 | w2 | w3 | w4 | uncert] (dev fp32, 16-byte aligned), groups as above (float ranges; begin a
 * multiple of 4, any end), a disabled group is left untouched.  Same arithmetic as nrt_adam_step. */
int nrt_adam_step_groups(float* param, float* grad, float* exp_avg, float* exp_avg_sq, const NrtAdamGroup* groups, int32_t n_groups,
                         int zero_grad, void* stream);
/* nrt_step_begin(map_counter, 1, ...) that also advances a second counter by one when it is given (the uncertainty grid's
 * Adam step count, on the iterations that step it): one launch opens the iteration. */
int nrt_iteration_begin(int32_t* map_counter_dev, int32_t* uncert_counter_dev, uint64_t seed, float* rand6_dev, void* stream);

/* ---- device-resident ray sampling ----------------------------------------------------------------
 * The host half of the mapping iteration on device-resident data.  Index lists are dev int64
 * arrays: either the reference's own `random.sample` draws (parity) or the output of nrt_sample_indices. */
/* get_camera_rays, type 'OpenGL' (tp/datasets/utils.py:24-57): dirs dev [H,W,3] = [(i-cx)/fx, -(j-cy)/fy, -1] */
int nrt_camera_rays(int32_t H, int32_t W, float fx, float fy, float cx, float cy, float* dirs, void* stream);
/* cat([direction, rgb, depth[...,None]], -1).reshape(-1, 7) (src/slam/coslam/coslam.py:290-291; keyframe.py:43-44) */
int nrt_pack_frame(const float* direction, const float* rgb, const float* depth, int64_t n_pixels, float* frame_rays, void* stream);
/* count (dev int32, overwritten) of pixels with 0 < depth <= depth_trunc (src/slam/coslam/coslam.py:319-321) */
int nrt_valid_depth_count(const float* frame_rays, int64_t n_pixels, float depth_trunc, int32_t* count, void* stream);
/* KeyFrameDatabaseNaruto.add_keyframe (src/slam/coslam/model/keyframe.py:21-60): slot[i] = frame_rays[idxs[i % n_idx]]
 * for i < rays_per_kf (the reference doubles the selected rows until there are enough of them).  n_valid_dev (optional, dev
 * int32): the valid-depth pixel count the indices were drawn from; when it is 0 the slot is left untouched, like the reference,
 * which attaches the frame id and returns before storing (rays.shape[1] == 0). */
int nrt_kf_store(const float* frame_rays, const int64_t* idxs, int64_t n_idx, int32_t rays_per_kf, const int32_t* n_valid_dev,
                 float* slot, void* stream);
/* random.sample(range(n), k): k distinct uniform indices (a keyed bijection of [0,n), seed-deterministic).  If n_dev != NULL
 * the population size is read from that device int32 (e.g. the valid-depth count), otherwise `n` is used. */
int nrt_sample_indices(int64_t n, const int32_t* n_dev, int64_t k, uint64_t seed, int64_t* out, void* stream);
/* KeyFrameDatabase.sample_global_rays (tp/model/keyframe.py:69-79) + the batch assembly and pose transform of global_BA
 * (src/slam/coslam/coslam.py:329-344).  kf_rays dev [num_kf*rays_per_kf,7]; frame_ids dev int64 [num_kf]; poses dev
 * [n_poses,4,4] whose last row is the current frame.  Outputs dev [n_global+n_cur, 3|3|3|1]. */
int nrt_assemble_rays(const float* kf_rays, const int64_t* frame_ids, int32_t rays_per_kf, int32_t keyframe_every,
                      const int64_t* idxs_global, int64_t n_global, const float* cur_rays, const int64_t* idx_cur, int64_t n_cur,
                      const float* poses, int32_t n_poses, float* rays_o, float* rays_d, float* target_s, float* target_d,
                      void* stream);
/* ActiveRaySampler.sample_rays (src/slam/coslam/active_ray_sampler.py:77-149): out = [the num_uncert_sample pool rays with the
 * LOWEST cached uncertainty (ascending pool order; ties at the threshold broken by lowest index) | rows 0 .. base-K | the
 * last ceil(n_cur/mul) rows], out_* dev [base + ceil(n_cur/mul), .].  vol_dims, bound_min: HOST arrays of 3.  chosen: dev
 * int32 [K] pool indices or NULL.  workspace: dev scratch of nrt_active_select_workspace(n_rays) bytes. */
int64_t nrt_active_select_workspace(int64_t n_rays);
int nrt_active_select(const float* rays_o, const float* rays_d, const float* target_s, const float* target_d, int64_t n_rays,
                      int64_t n_cur, const float* uncert_vol, const int32_t* vol_dims, const float* bound_min, int32_t base_sample_num,
                      int32_t num_uncert_sample, int32_t oversample_mul, float* out_o, float* out_d, float* out_s, float* out_t,
                      int32_t* chosen, void* workspace, void* stream);

/* ---- diagnostics ------------------------------------------------------------------------------- */
/* Tensor-core self-test (no reference counterpart): one CTA multiplies small fp32 matrices through the same
 * tcgen05 descriptors / TMEM read-back the MLP kernels use.  All pointers dev, row-major fp32.
 *   mode 0: d[128,n] = a[128,k] * b[n,k]^T  (k % 8 == 0, n % 16 == 0)
 *   mode 1: d[k,n]   = a[128 rows,k]^T * b[128 rows,n]  (weight-gradient form, k <= 128 valid output rows, passes = 1)
 *   mode 2: as mode 0 with the A operand staged in tensor memory
 * passes = 1 (plain TF32) or 3 (hi/lo split, ~fp32 accuracy).  d always has 128 rows. */
/* Profiling aid: with NRT_BWD_DEBUG=8 in the environment the backward kernel stamps clock64() at its phase boundaries
 * (per CTA: entry, prologue done, MLP loop done, scatter loop done, before flush, after flush, end); this copies the
 * [256][8] int64 table to host memory (synchronises).  A NEGATIVE `bytes` reads |bytes| of the forward kernel's table instead
 * (NRT_FWD_DEBUG=1: per CTA the cycles of sub-CTA 0 in prologue / ray staging / tiles / compositing / total / incl. statistics tail).
 * With bit 30 set in `bytes` (and NRT_PEER_DEBUG=1): the globaltimer stamps of nrt_adam_step_peers' phases (8 x u64). */
int nrt_debug_read(void* host_dst, int32_t bytes);
/* Profiling aid: a one-thread launch that writes the GPU's nanosecond %globaltimer to dst (dev u64) -- stage boundaries of a
 * captured iteration (MappingStep.trace_graph), which events cannot time inside a CUDA graph. */
int nrt_debug_stamp(uint64_t* dst_dev, void* stream);
int nrt_selftest_umma(int mode, const float* a, const float* b, int32_t k, int32_t n, int passes, float* d, void* stream);
/* Raw probe: a_img / b_img (dev) are copied verbatim into shared memory and multiplied as d[128,n] with the given
 * descriptor fields (bytes): leading / stride byte offsets, per-k-step start-address advance, MN-major flags. */
int nrt_selftest_umma_raw(const float* a_img, int32_t a_bytes, const float* b_img, int32_t b_bytes, int32_t n, int32_t ksteps,
                          int32_t a_mn, int32_t b_mn, int32_t a_lbo, int32_t a_sbo, int32_t a_kstep, int32_t b_lbo, int32_t b_sbo,
                          int32_t b_kstep, float* d, void* stream);

/* ---- planner hand-off on the device ------------------------------------------------------------
 * NarutoPlanner.uncertainty_aggregation_v2 (src/planner/naruto_planner.py:596-735) on device-resident volumes (the output
 * of nrt_map_volumes): for every goal-space candidate g and target voxel j, collections[g,j] = uncert[target j] if the
 * pair is inside the sensing range (min_dist < |g - j| < max_dist, voxels), the candidate is safe (one-voxel rim, SDF of
 * its 7-stencil >= safe_sdf) and the 30-sample segment between them has SDF > 0 everywhere, else 0; aggre[g] = sum_j.
 * uncert_vol / sdf_vol: dev [X,Y,Z]; dims = {X,Y,Z}; goal_pts: dev fp32 [n_goal,3] voxel coordinates (integers stored as
 * floats, as the reference's goal_space_pts); topk_vxl: dev fp32 [k,3] (the reference's argpartition draw, or any
 * selection); collections: dev [n_goal,k]; aggre: dev [n_goal]; n_valid (optional): dev int32, number of valid pairs
 * (0 = the reference's "invalid goal space"). */
int nrt_goal_aggregate(const float* uncert_vol, const float* sdf_vol, const int32_t* dims, const float* goal_pts, int64_t n_goal,
                       const float* topk_vxl, int32_t k, float min_dist, float max_dist, float safe_sdf, float* collections,
                       float* aggre, int32_t* n_valid, void* stream);

/* ---- iso-surface extraction ---------------------------------------------------------------------
 * mcubes.marching_cubes(volume, isovalue, truncation) of the reference
 * (third_parties/coslam/external/NumpyMarchingCubes/marching_cubes/src/marching_cubes.cpp:418-462 behind _mcubes.pyx:20-25;
 * called by save_mesh / save_uncert_mesh at src/slam/coslam/coslam_utils.py:145): dual-grid marching cubes with
 * truncation / jump thresholds, then an order-dependent vertex merge and face clean-up.  The O(n^3) part (corner means,
 * cell classification, prefix sum, vertex interpolation) runs on the device, the merge over the compacted triangle soup on
 * the host; vertices and faces equal the reference's, index for index.
 * volume: dev fp32 [nx,ny,nz] (the values the reference reads, which are float32 sweeps widened to double);
 * workspace: dev scratch of nrt_mc_workspace_bytes() bytes.  This is an export path: nrt_mc_extract synchronises the stream
 * and allocates the (result-sized) soup buffer itself.  The result handle owns host arrays: vertices double [V,3], faces
 * uint64 [F,3] (the reference's dtypes); copy them out, then free the handle. */
int64_t nrt_mc_workspace_bytes(int32_t nx, int32_t ny, int32_t nz);
int nrt_mc_extract(const float* volume, int32_t nx, int32_t ny, int32_t nz, float isovalue, float truncation, void* workspace,
                   void* stream, void** result);
int nrt_mc_result_sizes(void* result, int64_t* n_vertices, int64_t* n_faces);
int nrt_mc_result_copy(void* result, double* vertices, uint64_t* faces);
void nrt_mc_result_free(void* result);

/* ---- src/layers: ERP depth -> ERP radial distance ------------------------------------------------
 * ERPDepth2Dist.forward (src/layers/erp_conversions.py:288-354; called per simulator step at
 * src/simulator/habitat_simulator.py:143): 6 x E2P bilinear grid_sample -> depth2dist -> C2E nearest grid_sample, fused
 * into one pass over the output panorama.  erp_depth / erp_dist: dev [H,W].  The three static grids are what the
 * reference's constructor builds: c2e_grid dev [H,W,3] = C2E.grid (normalised x, y, face; src/layers/c2e.py:82-130),
 * face_coor dev [6,s,s,2] = the six E2P.coor_xy (order F R B L U D), face_rays dev [3,s*s] = K^-1 [u,v,1] of
 * Backprojection (src/layers/backprojection.py:31-82) with K = diag-ish(s/2). */
int nrt_erp_depth2dist(const float* erp_depth, int32_t H, int32_t W, const float* c2e_grid, const float* face_coor,
                       const float* face_rays, int32_t skybox_size, float* erp_dist, void* stream);
/* The same conversion without look-up grids: every grid entry above is a closed-form function of the output pixel and is
 * evaluated in the kernel (C2E.__init__ src/layers/c2e.py:82-130 with equirect_uvgrid / equirect_facetype
 * src/layers/c2e_utils.py:68-93; create_erp_coor src/layers/erp_conversions.py:184-229; Backprojection
 * src/layers/backprojection.py:31-82).  face_rot: HOST fp32 [6,3,3], row-vector frames of the faces F R B L U D
 * (p_panorama = [x, -y, 1] @ face_rot[f]); x_max = tan(fov/2) of a face (1 for the 90-degree skybox).  W % 4 == 0. */
int nrt_erp_depth2dist_analytic(const float* erp_depth, int32_t H, int32_t W, int32_t skybox_size, const float* face_rot,
                                float x_max, float* erp_dist, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NARUTO_B200_H */
