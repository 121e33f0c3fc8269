/*
 * hortimapping_b200 -- C ABI of the B200-native shape-completion / pose-estimation inner loop.
 *
 * The reference (PRBonn/HortiMapping) is pure Python/PyTorch and has no FFI of its own
 * (SURVEY.md section 2.1); the entry points below are what a ctypes binding of the reference's hot
 * path binds.  Each one names the reference interface it replaces (paths relative to the reference
 * root).  INTEGRATION.md shows the ctypes stub and how the reference's host scripts pick it up.
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch / C++ types.  All matrices row-major, fp32.
 *   - pointers prefixed d_ are DEVICE pointers on the context's device; h_ are HOST pointers.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default stream)
 *     except the *_host variants, which copy in, run, copy out and synchronise the stream.
 *   - return 0 on success, a negative HM_ERR_* code otherwise; hm_last_error() has the message.
 *     Nothing is thrown across the ABI.  Per-fruit data problems are NOT errors: they are reported in
 *     the status words (HM_STATUS_*), mirroring the reference's print-and-continue behaviour
 *     (wild_completion/optimizer.py:130-141).
 *   - a context is bound to one device and owns the split-precision copies of the decoder weights and grow-only scratch that
 *     every call reuses.  Calls of one context are therefore serialised on the GPU even across streams (a call issued on
 *     another stream than the previous one first waits, device-side, for that call's work); use one context per stream for
 *     concurrency.  A context is not thread-safe: calls from several host threads need external locking.
 */
#ifndef HORTIMAPPING_B200_H
#define HORTIMAPPING_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HM_LATENT 32            /* specs.json "CodeLength"                               */
#define HM_IN 35                /* latent ++ xyz (deep_sdf_decoder.py:27)                  */
#define HM_HIDDEN 512           /* specs.json NetworkSpecs.dims                            */
#define HM_LAYERS 9             /* lin0 .. lin8                                            */
#define HM_MAX_POSE 7           /* Sim(3): translation(3) rotation(3) scale(1)             */
#define HM_MAX_EST (HM_MAX_POSE + HM_LATENT)

#define HM_OK 0
#define HM_ERR_INVALID (-1)     /* bad argument / unsupported architecture                 */
#define HM_ERR_CUDA (-2)        /* a CUDA runtime call failed                              */
#define HM_ERR_NOMEM (-3)

/* decoder engines */
#define HM_ENGINE_TC 0          /* tcgen05 tensor-core path, split-fp16 operands (default) */
#define HM_ENGINE_SIMT 1        /* fp32 CUDA-core kernels (validation / calibration path)  */

/* per-fruit status bits written by the optimisers */
#define HM_STATUS_CONV_GRADIENT 0x01   /* optimizer.py:276  max|b| < epsilon_g             */
#define HM_STATUS_CONV_CODE 0x02       /* optimizer.py:280                                  */
#define HM_STATUS_CONV_POSE 0x04       /* optimizer.py:285                                  */
#define HM_STATUS_MAX_ITER 0x08        /* optimizer.py:289                                  */
#define HM_STATUS_FRAME_SKIPPED 0x10   /* optimizer.py:130-132 "This frame is not valid"    */
#define HM_STATUS_SUBMAP_INVALID 0x20  /* optimizer.py:139-141 "This submap is not valid"   */
#define HM_STATUS_F16_SATURATED 0x40   /* a row of THIS fruit left the calibrated fp16 range of the TC engine */
#define HM_STATUS_SOLVE_FAILED 0x80    /* singular / non-finite LM system: the fruit's state was left untouched  */

typedef struct hm_context hm_context;

/* Folded decoder weights (weight-norm already applied: W = g * v / ||v||, deep_sdf_decoder.py:49-54),
 * HOST pointers, layer l is [out_dim[l]][in_dim[l]] row-major.  Only the architecture the reference
 * ships is accepted: 8 x 512 hidden, latent 32, latent_in = [4] (lin3 has 512-35 outputs), 1 output. */
typedef struct hm_decoder_desc {
  int32_t n_layers;
  int32_t latent_size;
  int32_t latent_in_layer;
  int32_t in_dim[HM_LAYERS];
  int32_t out_dim[HM_LAYERS];
  const float* weight[HM_LAYERS];
  const float* bias[HM_LAYERS];
} hm_decoder_desc;

/* POD mirror of cfg['opt'] as read at wild_completion/optimizer.py:31-53 (+ loss.py:11 defaults). */
typedef struct hm_opt_params {
  int32_t max_iter;           /* converge.max_iter                                          */
  int32_t n_depth_samples;    /* render.n_sample_on_ray (<= 64)                             */
  int32_t log_sdf_occ;        /* render.log_sdf_occ                                         */
  int32_t occlusion_on;       /* render.occlusion_on                                        */
  int32_t lm_on, lm_eye;
  int32_t robust_iter;
  int32_t scale_on;
  int32_t min_valid_sample;   /* loss.py:11 default 100                                     */
  int32_t iter_offset;        /* test hook: loop index starts here (0 in the reference)     */
  /* real-valued settings are Python floats (doubles) in the reference and meet fp32 tensors only
   * at the point of use; they are carried as doubles so that the casts happen where torch's do */
  double epsilon_g, epsilon_c, epsilon_t, epsilon_r, epsilon_s;
  double occ_cutoff_m;        /* render.occ_cutoff_m                                        */
  double w_recon, w_depth, w_mask, w_codereg;
  double lm_lambda_0;
  double robust_th_recon;     /* recon.robust_th_m                                          */
  double robust_th_depth;     /* render.robust_th_m                                         */
  double s_damp;
  double occlusion_th;        /* loss.py:11 default 0.03                                    */
  double min_grad_thre;       /* loss.py:11 default 1e-6                                    */
} hm_opt_params;

/* A batch of independent fruits.  Offsets are HOST arrays (the host knows the sizes; they set launch
 * and workspace sizes); bulk data are DEVICE arrays.  Frames must already be sub-sampled the way
 * optimizer.py:77-78 does (np.linspace over the matched frames). */
typedef struct hm_fruit_batch {
  int32_t n_fruits;
  float* d_latents;                /* [n_fruits][32]  in/out (optimizer.py:248 updates in place)   */
  float* d_T_ow;                   /* [n_fruits][16]  in/out                                       */
  const float* d_points_w;         /* [n_points][3]   world-frame surface points                   */
  const int64_t* h_point_offsets;  /* [n_fruits+1]                                                 */
  const int32_t* h_frame_offsets;  /* [n_fruits+1]    NULL for shape-only optimisation             */
  const float* d_T_wc;             /* [n_frames][16]                                               */
  const int64_t* h_ray_offsets;    /* [n_frames+1]                                                 */
  const int32_t* h_n_fg;           /* [n_frames]      the first n_fg rays of a frame are foreground */
  const float* d_rays;             /* [n_rays][3]     camera-frame directions, z = 1               */
  const float* d_depth_obs;        /* [n_rays]        observed z-depth (0 = none)                  */
  const float* h_cube_radius;      /* [n_fruits]                                                   */
  const uint8_t* h_pose_known;     /* [n_fruits]      optimizer.py:237-238                         */
  int32_t* d_iter_count;           /* [n_fruits] out                                               */
  int32_t* d_status;               /* [n_fruits] out  HM_STATUS_* bits                             */
} hm_fruit_batch;

/* Counters, cumulative since hm_create (for roofline accounting).  The row and tile counts are EXACT: the decoder
 * kernels add the number of rows they actually evaluated (the device-side count of compacted ray samples / in-band samples, not
 * the launch upper bound) to device counters that hm_get_counters reads back. */
typedef struct hm_counters {
  int64_t rows_forward;       /* decoder rows evaluated forward-only                                            */
  int64_t rows_jacobian;      /* decoder rows evaluated forward + input gradient                                */
  int64_t kernel_launches;    /* kernels of this library launched                                               */
  int64_t iterations;         /* LM iterations launched (max over fruits)                                       */
  int64_t decoder_launches;   /* decoder kernel launches timed while profiling was enabled                      */
  double decoder_ms;          /* sum of their CUDA-event durations (ms)                                         */
  int64_t forward_launches;   /* ... of which forward-only launches                                             */
  double forward_ms;
  int64_t jacobian_launches;  /* ... and forward + input-gradient launches                                      */
  double jacobian_ms;
  int64_t tiles_forward;      /* 64-row tiles processed by forward-only / forward+gradient launches             */
  int64_t tiles_jacobian;
  int64_t tiles_redone_forward;  /* ... of which contradicted the sparse plan and were re-evaluated with the full  */
  int64_t tiles_redone_jacobian; /*     plan (hm_set_sparse_plan; 0 for the shipped models on calibrated rows)     */
  int64_t rows_backward;      /* decoder rows evaluated gradient-only from stored ReLU masks (hm_set_mask_reuse)  */
  int64_t tiles_backward;
  int64_t tiles_redone_backward;
  int64_t backward_launches;  /* timed gradient-only launches and their CUDA-event durations                     */
  double backward_ms;
} hm_counters;

const char* hm_last_error(void);
int hm_version(void);

/* deepsdf/deep_sdf/workspace.py:203-225 config_decoder: build the decoder on `device`. */
int hm_create(hm_context** out, int device, const hm_decoder_desc* dec);
void hm_destroy(hm_context* ctx);
int hm_set_engine(hm_context* ctx, int engine);
int hm_get_engine(const hm_context* ctx);
/* Tensor-core engine: the sparse plan (on by default).  hm_calibrate records which hidden units were ever alive on the
 * calibration rows; the engine orders the units so that those come first and drops every MMA whose A operand is then an
 * all-zero 64-wide chunk of the activations (both shipped models: lin3 is dead outright, 12 .. 320 of 512 units alive elsewhere).
 * The assumption is checked per 64-row tile from the ReLU bits; a tile that contradicts it is re-evaluated with the full plan
 * by a second launch (hm_counters.tiles_redone_*).  Dropped products are exact zeros: results are bit-identical with the plan
 * off; the switch exists for that test and for measurements. */
int hm_set_sparse_plan(hm_context* ctx, int on);
/* What the tensor-core engine ISSUES per decoder row under the current calibration (measurement aid: the roofline's algorithmic
 * FLOPs are the reference's dense count, deep_sdf_decoder.py:75-110; this is the work actually sent to the tensor cores):
 * h_out[0] / [1] = tensor-core FLOP per row of a forward / forward + gradient evaluation with the sparse plan, [2] / [3] = the
 * same with the full plan (three fp16 products per fp32 product, padding included), [4 .. 11] = 64-wide chunks of h_0 .. h_7 the
 * sparse plan treats as possibly non-zero (8 = no assumption).  h_out has 12 entries. */
int hm_plan_info(const hm_context* ctx, double* h_out);
/* Joint loop (hm_optimize_joint), tensor-core engine: reuse of the forward pass (on by default).  The reference evaluates the
 * in-band ray samples twice per iteration -- once among all in-sphere samples (loss.py:47-49, no grad) and once more with autograd
 * for the Jacobians (loss.py:185-215).  With the switch on, the forward launch stores every row's ReLU bits (512 B per row) and the
 * gradient of the in-band samples comes from a gradient-only launch that starts from those bits and the SDF values already
 * computed; only the observed points (loss.py:219-243) still need forward + gradient.  Same arithmetic on the same operands:
 * results are bit-identical with the switch off; it exists for that test and for measurements. */
int hm_set_mask_reuse(hm_context* ctx, int on);
/* Choose the power-of-two fp16 operand scales of the TC engine from sample rows [n][35] (device). */
int hm_calibrate(hm_context* ctx, const float* d_rows, int64_t n, void* stream);
/* With profiling enabled every decoder kernel launch is bracketed by CUDA events on its stream; hm_get_counters synchronises the
 * device, adds the recorded durations and reads the device-side row / tile counters. */
int hm_get_counters(hm_context* ctx, hm_counters* out);
/* Number of saturation events of the tensor-core decoder since the last call (one per 64-row tile and thread that saw an operand
 * leave the calibrated fp16 range: the conversion saturates, the result is then NOT fp32-grade -- re-run hm_calibrate on
 * representative rows).  Counts decoder calls and optimiser launches alike; synchronises the device and resets the count.  The
 * optimisers additionally report the condition per FRUIT as HM_STATUS_F16_SATURATED. */
int hm_saturation_count(hm_context* ctx, int64_t* h_count);
int hm_profile_enable(hm_context* ctx, int on);

/* wild_completion/utils.py:144-172 decode_sdf: sdf[i] = f(latent, xyz[i]). */
int hm_sdf_forward(hm_context* ctx, const float* d_latent, const float* d_xyz, int64_t n, float* d_sdf, void* stream);
/* deepsdf/networks/deep_sdf_decoder.py:75-110 Decoder.forward on arbitrary rows [n][35]. */
int hm_sdf_forward_rows(hm_context* ctx, const float* d_rows, int64_t n, float* d_sdf, void* stream);
/* wild_completion/utils.py:175-193 get_batch_sdf_jacobian: sdf[n], jac[n][35] = d sdf / d[latent, xyz]. */
int hm_sdf_jacobian(hm_context* ctx, const float* d_latent, const float* d_xyz, int64_t n, float* d_sdf, float* d_jac, void* stream);
int hm_sdf_jacobian_rows(hm_context* ctx, const float* d_rows, int64_t n, float* d_sdf, float* d_jac, void* stream);

/* wild_completion/utils.py:542-562 create_voxel_grid(vol_dim) * cube_radius -> xyz[vol_dim^3][3]. */
int hm_voxel_grid(hm_context* ctx, int32_t vol_dim, float cube_radius, float* d_xyz, void* stream);
/* wild_completion/mesher.py:12-18: SDF of `latent` on that grid -> sdf[vol_dim^3] (C order). */
int hm_sdf_grid(hm_context* ctx, const float* d_latent, int32_t vol_dim, float cube_radius, float* d_sdf, void* stream);

/* wild_completion/utils.py:565-588 convert_sdf_voxels_to_mesh, on the device: zero level set of d_sdf [n][n][n] (C order) by
 * marching tetrahedra (the reference calls skimage.measure.marching_cubes on the host; mesh parity is pinned at Chamfer
 * level).  The mesh stays in the context; the counts come back on the host (synchronises the stream). */
int hm_isosurface(hm_context* ctx, const float* d_sdf, int32_t n, double level, double spacing, int64_t* h_n_verts,
                  int64_t* h_n_faces, void* stream);
/* Copies that mesh to d_verts [n_verts][3] fp32 and d_faces [n_faces][3] int32; apply_affine != 0 maps the vertices to
 * (v - 1) * cube_radius as utils.py:583-585 does. */
int hm_isosurface_fetch(hm_context* ctx, float* d_verts, int32_t* d_faces, int32_t apply_affine, double cube_radius, void* stream);

/* wild_completion/utils.py:39-109 get_render_data, device side (ctx may be NULL: current device).
 * hm_frame_id_bboxes (:48-59): ONE pass over a frame's submap-id image d_id_img [h][w] (int32) and depth image d_depth [h][w]:
 *   d_table [n_ids][5] = (count, min_v, max_v, min_u, max_u) of the pixels with that id and depth > 0 (count 0: min = INT32_MAX,
 *   max = -1), for every id < n_ids (<= 1024) at once.
 * hm_crop_candidates (:66-73, :81-84): the crop grid rows d_hh [crop_h] x columns d_ww [crop_w] in row-major order; background
 *   candidates (id != submap_id) and foreground candidates (id == submap_id and depth > 0) as [u, v] pixels + depths, in the
 *   reference's order; d_counts [2] = (n_bg, n_fg).  Output buffers hold crop_h * crop_w entries.
 * hm_gather_rays (:23-38 get_rays, :74-76, :85-87): for the k selected candidates d_sel (NULL = the first k) the ray directions
 *   float32(([u, v, 1] * invK).sum(-1)) (fp64 arithmetic, h_invK [9] row-major on the host), depths and pixels. */
int hm_frame_id_bboxes(hm_context* ctx, const int32_t* d_id_img, const float* d_depth, int32_t h, int32_t w, int32_t n_ids,
                       int32_t* d_table, void* stream);
int hm_crop_candidates(hm_context* ctx, const int32_t* d_id_img, const float* d_depth, int32_t h, int32_t w, int32_t submap_id,
                       const int32_t* d_hh, int32_t crop_h, const int32_t* d_ww, int32_t crop_w, int32_t* d_pix_bg, float* d_depth_bg,
                       int32_t* d_pix_fg, float* d_depth_fg, int32_t* d_counts, void* stream);
int hm_gather_rays(hm_context* ctx, const int32_t* d_pix, const float* d_depth, const int64_t* d_sel, int64_t k, const double* h_invK,
                   float* d_rays, float* d_depth_out, int32_t* d_pix_out, void* stream);

/* metrics_3d/chamfer_distance.py:23-24, metrics_3d/precision_recall.py:33-36 (open3d compute_point_cloud_distance):
 * d_dist[i] = distance from d_query[i] to its nearest neighbour in d_target; fp64, [n][3] row-major, exact. */
int hm_nn_distance(hm_context* ctx, const double* d_query, int64_t n_query, const double* d_target, int64_t n_target,
                   double* d_dist, void* stream);

/* wild_completion/utils.py:408-419 clean_pcd, device side: DBSCAN labels of d_points [n][3] (fp64) exactly as open3d's
 * PointCloud::cluster_dbscan(eps, min_points) assigns them (neighbourhood |p - q|^2 < eps^2 including p itself; clusters numbered
 * by their smallest core index; border points take the smallest cluster number among their core neighbours; noise = -1), computed
 * order-independently in parallel.  d_labels [n] int32; h_n_clusters may be NULL.  ctx may be NULL (current device).
 * Synchronises the stream. */
int hm_dbscan(hm_context* ctx, const double* d_points, int64_t n, double eps, int32_t min_points, int32_t* d_labels,
              int32_t* h_n_clusters, void* stream);
/* wild_completion/utils.py:426-427 (get_axis_aligned_bounding_box): component-wise min / max of d_points [n][3] (fp64). */
int hm_cloud_bounds(hm_context* ctx, const double* d_points, int64_t n, double* h_min3, double* h_max3, void* stream);
/* wild_completion/utils.py:447-455: the points of d_points inside the box [h_box_min3, h_box_max3] (inclusive, open3d's
 * AxisAlignedBoundingBox crop): their count and the mean of (p - h_center3). */
int hm_crop_mean_offset(hm_context* ctx, const double* d_points, int64_t n, const double* h_box_min3, const double* h_box_max3,
                        const double* h_center3, int64_t* h_count, double* h_mean3, void* stream);

/* wild_completion/loss.py:219-243 compute_sdf_loss: res[n], J_pose[n][pose_dim], J_code[n][32]. */
int hm_sdf_loss(hm_context* ctx, const float* d_latent, const float* d_pts_obj, int64_t n, int32_t scale_on,
                float* d_res, float* d_J_pose, float* d_J_code, void* stream);

/* wild_completion/loss.py:8-217 compute_render_loss for ONE frame, same inputs as the reference call at
 * optimizer.py:116-118.  Outputs are per INPUT ray (n_rays rows): d_ray_valid[r] != 0 marks the rays the
 * reference returns (ascending order); J rows are [pose_dim + 32] wide (pose first); h_T_oc (16 floats)
 * and h_depths (n_depth_samples floats) are HOST arrays; h_n_valid_samples receives the in-sphere
 * sample count (the reference returns None when it is < min_valid_sample; the outputs are then
 * all-invalid).  Synchronises the stream. */
int hm_render_loss(hm_context* ctx, const hm_opt_params* p, const float* d_latent, const float* d_rays, int32_t n_rays,
                   int32_t n_fg, const float* d_depth_obs, const float* h_T_oc, const float* h_depths,
                   float bbx_radius, int32_t* d_ray_valid, float* d_res_d, float* d_J_d, float* d_res_m,
                   float* d_J_m, int32_t* h_n_valid_samples, void* stream);

/* wild_completion/optimizer.py:306-429 shape_opt_deepsdf for a batch of fruits (latent only). */
int hm_optimize_shape(hm_context* ctx, const hm_opt_params* p, const hm_fruit_batch* batch, void* stream);
/* wild_completion/optimizer.py:28-302 shape_pose_joint_opt for a batch of fruits. */
int hm_optimize_joint(hm_context* ctx, const hm_opt_params* p, const hm_fruit_batch* batch, void* stream);
/* Diagnostics / test hook: H [n_fruits][est*est], b [n_fruits][est], dx [n_fruits][est] of the LAST iteration run, for the first
 * n_fruits fruits of the last optimise call (an error if it had fewer; est = pose_dim + 32, pose_dim = 0 for hm_optimize_shape).
 * Device pointers, may be NULL. */
int hm_get_last_system(hm_context* ctx, int32_t n_fruits, float* d_H, float* d_b, float* d_dx, void* stream);

/* Same optimisers with every pointer of `batch` a HOST pointer: copies in, runs, copies the results
 * (latents, T_ow, iter_count, status) back and synchronises.  This is the call bench.py's e2e times. */
int hm_optimize_shape_host(hm_context* ctx, const hm_opt_params* p, const hm_fruit_batch* host_batch);
int hm_optimize_joint_host(hm_context* ctx, const hm_opt_params* p, const hm_fruit_batch* host_batch);

#ifdef __cplusplus
}
#endif
#endif /* HORTIMAPPING_B200_H */
