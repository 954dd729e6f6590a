/*
 * rcvvote.h -- C ABI of librcvvote.so: the B200-native (sm_100a) radial keypoint-voting path.
 *
 * The reference (aaronWool/rcvpose) has no plugin/FFI layer: the path is three Python callables
 * (SURVEY.md section 8b).  This header is the boundary a maintainer binds instead; each entry point
 * names the reference interface it replaces (file:line in the reference repository).
 *
 * Conventions
 *   - Plain C: pointers, sizes, int status codes.  No C++ exceptions cross the boundary.
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).
 *   - Unless a function name ends in `_host`, every data pointer is a DEVICE pointer on the
 *     context's device and the call is asynchronous on `stream` (outputs are valid once the
 *     stream has completed).  `_host` variants take HOST pointers, perform the host<->device
 *     copies on `stream` and synchronise it before returning.
 *   - The caller owns every input/output buffer and the stream; the context owns scratch.
 *     A context is not re-entrant; use one context per (device, stream).
 *   - Return value: RCV_OK or a negative RCV_E_* code; rcv_last_error() gives the text.
 *     Data-dependent conditions are reported per item in `status[]` (RCV_ST_*), never as UB.
 */
#ifndef RCVVOTE_H_
#define RCVVOTE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RCV_ABI_VERSION 2

/* call status */
#define RCV_OK 0
#define RCV_E_INVALID (-1)   /* bad argument */
#define RCV_E_CUDA (-2)      /* CUDA runtime error (see rcv_last_error) */
#define RCV_E_CAPACITY (-3)  /* request exceeds the capacities fixed at rcv_create */
#define RCV_E_NOGPU (-4)     /* no usable sm_100 device: there is no CPU fallback */

/* per-item status bits (status[] outputs) */
#define RCV_ST_OK 0
#define RCV_ST_EMPTY_MASK 1      /* no surviving pixel/point: the reference raises ValueError (AccumulatorSpace.py:390) */
#define RCV_ST_BAD_GRID 2        /* D <= 0 (np.zeros would raise) */
#define RCV_ST_D_EXCEEDS_CAP 4   /* D > cfg.max_grid */
#define RCV_ST_POINT_OVERFLOW 8  /* compacted points exceed cfg.max_points_total */
#define RCV_ST_UNIT_OVERFLOW 16  /* tile work list exceeds capacity */

/* grid policy: which Accumulator_3D prelude is reproduced */
#define RCV_POLICY_LM 0      /* AccumulatorSpace.py:373-419: D = int(max) + int(rmax); centre = (idx+mean+0.5)*unit */
#define RCV_POLICY_YCBGEN 1  /* 3DRadius_ycb.py:113-161:    D = int(max) + 1;         centre = (idx+mean)*unit+0.5 */

/* element types */
#define RCV_F32 0
#define RCV_F64 1
#define RCV_U16 2

/* mask rule flags for rcv_vote_frames (AccumulatorSpace.py:603-618, 837-851, 1049-1053) */
#define RCV_MASK_RADIUS_NONZERO 1   /* keep pixel iff radius != 0   (LM  npy branch :613) */
#define RCV_MASK_RADIUS_POSITIVE 2  /* keep pixel iff radius > 0    (LMO npy branch :850) */
#define RCV_MASK_SEM_GT 4           /* keep pixel iff sem >  sem_threshold (LM/YCB ckpt :603,:1049) */
#define RCV_MASK_SEM_GE 8           /* keep pixel iff sem >= sem_threshold (LMO ckpt :837) */
#define RCV_MASK_MAX_RADIUS 16      /* keep pixel iff radius <= max_radii[k] (:604,:613,:838,:849) */

typedef struct rcv_ctx rcv_ctx;

typedef struct rcv_config {
  int abi_version;             /* RCV_ABI_VERSION */
  int max_items;               /* max (frame,keypoint) items per call */
  long long max_points_total;  /* pool for compacted points summed over the items of one call */
  int max_grid;                /* largest accumulator side D accepted */
  int max_units;               /* tile work-list capacity (0 = derive from max_items/max_grid) */
  /* Scratch that depends on sizes only a call knows.  Given here, rcv_create allocates it and no entry point allocates on the
   * hot call; left 0, the first call of a (larger) size allocates it after a stream synchronisation. */
  long long image_pixels;      /* H*W of the frames entry points: survival-bit scratch for max_items items (0 = on first use) */
  int max_model_points;        /* CAD points of rcv_add_metric_batch / rcv_icp_batch, for up to max_items frames (0 = on first use) */
  int head_items;              /* items whose radius planes rcv_head_vote_frames keeps in the context (0 = on first use) */
} rcv_config;

/* Parameters of Accumulator_3D that the reference hard-codes (AccumulatorSpace.py:374,388). */
typedef struct rcv_vote_params {
  double acc_unit;      /* 5   : voxel size in mm (:374) */
  double radius_scale;  /* 100 : radius unit -> mm (decimetre maps, :388); 1000 for metre radii (3DRadius_ycb.py:124) */
  int grid_policy;      /* RCV_POLICY_* */
  int radius_dtype;     /* RCV_F32 or RCV_F64: radius arithmetic is done in the radius' own dtype (:388) */
} rcv_vote_params;

typedef struct rcv_frame_params {
  int height, width;    /* 480, 640 */
  int depth_dtype;      /* RCV_U16 (LINEMOD .dpt, :482-490), RCV_F32 or RCV_F64 (LMO png, :833) */
  double depth_div;     /* depth is divided by this first (YCB factor_depth, :1051-1052); 1 = none */
  double xyz_div;       /* back-projected cloud is divided by this (mm -> m, :619); 1 = none (YCB) */
  int mask_flags;       /* RCV_MASK_* */
  float sem_threshold;  /* 0.8 (LM, YCB) or 0.5 (LMO) */
  int k_stride;         /* doubles between per-frame intrinsics (9), or 0 for one shared K */
  int max_radii_stride; /* doubles between per-frame max_radii rows (n_kpts), or 0 for one shared row */
} rcv_frame_params;

/* ---- lifetime ------------------------------------------------------------------------------ */
int rcv_create(int device, const rcv_config* cfg, rcv_ctx** out);
void rcv_destroy(rcv_ctx* ctx);
const char* rcv_last_error(const rcv_ctx* ctx); /* ctx may be NULL: error of the last failed rcv_create */
int rcv_abi_version(void);

/* ---- rgbd_to_point_cloud(K, depth)  -- AccumulatorSpace.py:77-85 ------------------------------
 * depth: (H,W) of `depth_dtype`; K: 9 doubles row-major (DEVICE).  Writes the (N,3) float64 cloud in
 * row-major pixel order to xyz_out (capacity xyz_capacity points) and N to *n_out (device int).  */
int rcv_backproject(rcv_ctx* ctx, const double* K, const void* depth, int depth_dtype, int height, int width,
                    double* xyz_out, long long xyz_capacity, int* n_out, void* stream);

/* ---- Accumulator_3D(xyz, radial_list)  -- AccumulatorSpace.py:373-419 ---------------------------
 * Batched over n_items independent point clouds stored back to back:
 *   xyz     (sum N, 3) float64, metres            radii  (sum N) of params->radius_dtype
 *   item_offsets (n_items+1) int64 on the DEVICE: item b owns points [off[b], off[b+1]).
 * Outputs per item (device): centre_mm[3] float64 (row 0 of the reference's result), peak = vote count of
 * the winning voxel, votes = total votes cast inside the grid (int64), grid = D, zero_boundary = zb,
 * status = RCV_ST_*.  Any output pointer except centre_mm/status may be NULL.
 * volume_out (optional, n_items must be 1): int32 [D][D][D] C-order dump of the vote volume
 * (the reference's VoteMap_3D), for parity checks; volume_capacity in int32 elements. */
int rcv_vote_points(rcv_ctx* ctx, const double* xyz, const void* radii, const long long* item_offsets, int n_items,
                    const rcv_vote_params* params, double* centre_mm, int* peak, long long* votes, int* grid,
                    int* zero_boundary, int* status, int32_t* volume_out, long long volume_capacity, void* stream);

/* ---- the per-frame loop body of estimate_6d_pose_*  -- AccumulatorSpace.py:601-637 (LM), 833-873 (LMO),
 *      1048-1067 (YCB): mask rule + rgbd_to_point_cloud + radial_list gather + Accumulator_3D, fused, for
 *      n_frames x n_kpts items.
 *   depth   [n_frames][H][W] depth_dtype           radius [n_frames][n_kpts][H][W] float32 (decimetres)
 *   sem     [n_frames][n_kpts][H][W] float32 or NULL
 *   K       [n_frames or 1][9] float64             max_radii [n_frames or 1][n_kpts] float64 or NULL
 * Outputs are indexed [frame][kpt]: centre_mm [..][3] float64, peak, votes, n_points, grid, status. */
int rcv_vote_frames(rcv_ctx* ctx, int n_frames, int n_kpts, const void* depth, const float* radius, const float* sem,
                    const double* K, const double* max_radii, const rcv_frame_params* fp, const rcv_vote_params* vp,
                    double* centre_mm, int* peak, long long* votes, int* n_points, int* grid, int* status, void* stream);

/* Same, HOST buffers in and out: copies inputs host->device and results device->host on `stream`,
 * in chunks that overlap transfer with voting, and synchronises before returning.  For full
 * bandwidth the host buffers should be page-locked (cudaHostAlloc / torch pin_memory).
 * Only the rows between the first and the last non-zero depth row of each frame are copied (every mask rule needs
 * depth != 0, so the rest cannot contribute; results are bit-identical); rcv_last_h2d_bytes() reports what was moved. */
int rcv_vote_frames_host(rcv_ctx* ctx, int n_frames, int n_kpts, const void* depth, const float* radius, const float* sem,
                         const double* K, const double* max_radii, const rcv_frame_params* fp, const rcv_vote_params* vp,
                         double* centre_mm, int* peak, long long* votes, int* n_points, int* grid, int* status,
                         int frames_per_chunk, void* stream);

/* ---- argwhere(V == V.max())[0]  -- AccumulatorSpace.py:406 ---------------------------------------
 * Peak search over an int32 [D][D][D] volume: first maximum in C order.  idx_out[3], max_out (device). */
int rcv_argmax_volume(rcv_ctx* ctx, const int32_t* volume, int grid, int* idx_out, int* max_out, void* stream);

/* ---- HornPoseFitting.lmshorn(P1, P2, n, A)  -- util/horn.py:75-181 --------------------------------
 * Batched closed-form absolute orientation: model [n_frames or 1][n][3] (model_stride = 3n or 0),
 * est [n_frames][n][3] float64 -> RT [n_frames][4][4] float64 with est ~= R*model + T. */
int rcv_horn_batch(rcv_ctx* ctx, const double* model, long long model_stride, const double* est, int n, int n_frames,
                   double* RT, void* stream);
int rcv_horn_batch_host(rcv_ctx* ctx, const double* model, long long model_stride, const double* est, int n, int n_frames,
                        double* RT, void* stream);

/* ---- ADD(-S) distance before ICP  -- AccumulatorSpace.py:664-702 (LM), 897-927 (LMO), 1119-1150 (YCB) -----------
 * The step right after the path: the CAD points are transformed by the estimated pose and
 * by the ground-truth pose (project(), :64-75: xyz @ R^T + t) and, for every ground-truth point, the distance to the
 * NEAREST estimated point is taken (open3d compute_point_cloud_distance, :688/:692); the reference thresholds the
 * MEAN of these distances, or their MINIMUM for the symmetric classes.  Exact float64 nearest neighbour (through a grid over the
 * model for models of 1,024 points and more; bit-identical to the all-pairs search, RCV_ADD_BRUTE=1).
 *   model_mm [n_model][3] float64 (one CAD model shared by the frames), RT_est / RT_gt [n_frames][4][4] float64
 *   row-major (rows 0..2 used; translations in the unit of model_mm), outputs mean_out / min_out [n_frames] float64. */
int rcv_add_metric_batch(rcv_ctx* ctx, const double* model_mm, int n_model, const double* RT_est, const double* RT_gt, int n_frames,
                         double* mean_out, double* min_out, void* stream);

/* ---- scene cloud of a frame (the ICP target)  -- AccumulatorSpace.py:620-625 (LM), :863-868 (LMO), :1070-1075 (YCB) ----
 * The reference appends, keypoint after keypoint, the masked cloud's points that were not seen before (`xyz_mm_icp`, an
 * O(N^2) Python loop).  Every cloud of a frame is a back-projection of the same depth map, so the union is
 * rgbd_to_point_cloud(K, depth * (mask_1 | ... | mask_Kp)) (:77-85), times `scale` (1 for LM/LMO whose clouds are in mm,
 * 1000 for YCB's xyz_icp*1000, :1154).  Same inputs and mask rules as rcv_vote_frames.  Points come out frame after
 * frame in row-major pixel order: xyz_out [offsets_out[f], offsets_out[f+1]) x 3 float64; offsets_out has n_frames + 1
 * entries; a frame that does not fit xyz_capacity (points) gets an empty range and RCV_ST_POINT_OVERFLOW in status_out
 * (may be NULL), an empty union RCV_ST_EMPTY_MASK. */
int rcv_scene_clouds(rcv_ctx* ctx, int n_frames, int n_kpts, const void* depth, const float* radius, const float* sem,
                     const double* K, const double* max_radii, const rcv_frame_params* fp, double scale, double* xyz_out,
                     long long xyz_capacity, long long* offsets_out, int* status_out, void* stream);

/* The same scene clouds for the frames of the MOST RECENT rcv_vote_frames / rcv_head_vote_frames call on this context, from the
 * survival bits that call left behind (their OR over the keypoints): no map is read again -- with the fused head the seg
 * plane never existed in memory.  Consumes the bits; RCV_E_INVALID if the last frames call had another shape. */
int rcv_scene_clouds_last(rcv_ctx* ctx, int n_frames, int n_kpts, const void* depth, const double* K, const rcv_frame_params* fp,
                          double scale, double* xyz_out, long long xyz_capacity, long long* offsets_out, int* status_out, void* stream);

/* ---- point-to-point ICP refinement of the Horn pose  -- AccumulatorSpace.py:704-718 (LM), :929-950 (LMO), :1152-1180 (YCB) ----
 * Replaces open3d 0.14.1 (rcvpose.yml:176; third-party, not in the reference tree)
 *   registration_icp(source = CAD model, target = scene, max_correspondence_distance, init,
 *                    TransformationEstimationPointToPoint(), ICPConvergenceCriteria(relative_fitness, relative_rmse, max_iteration))
 * for n_frames frames at once (see csrc/refine.cu for the restated algorithm).  Exact float64 nearest neighbour within max_dist
 * (a uniform grid over each frame's scene, built once per call; bit-identical to the all-pairs search, RCV_ICP_BRUTE=1).
 *   model [n_model][3] float64 (shared by the frames), scene [*][3] float64 with frame f owning
 *   [scene_offsets[f], scene_offsets[f+1]) (the layout rcv_scene_clouds writes), RT_init [n_frames][4][4] row-major,
 *   max_dist [n_frames] (the reference passes the ADD(-S) distance before ICP), defaults of open3d: max_iter 30,
 *   rel_fitness = rel_rmse = 1e-6.  Outputs: RT_out [n_frames][4][4] (reg.transformation), fitness_out / rmse_out
 *   [n_frames] (reg.fitness, reg.inlier_rmse), iters_out [n_frames] (updates applied).  Stream-ordered, except that the call
 *   reads scene_offsets[n_frames] (8 bytes) to size the grid scratch, and for max_iter > 32 looks at a device counter every 32
 *   iterations and stops once every frame has converged. */
int rcv_icp_batch(rcv_ctx* ctx, const double* model, int n_model, const double* scene, const long long* scene_offsets,
                  const double* RT_init, const double* max_dist, int n_frames, int max_iter, double rel_fitness, double rel_rmse,
                  double* RT_out, double* fitness_out, double* rmse_out, int* iters_out, void* stream);

/* ---- conv8 of the radius-map producer  -- models/fcnresnet.py:118 (definition), :187-189 (use) -----------
 * out[b][n][p] = bias[n] + sum_k bf16(weight[n][k]) * up[b][k][p],  n = 0 (seg), 1 (radial), k = 0..31.
 * The 1x1 head as a tcgen05 tensor-core kernel (bf16 operands, fp32 accumulation in tensor memory).
 *   up      [n_images][32][hw] bfloat16 (the NCHW output of conv7 + BN + ReLU), 16-byte aligned, hw % 8 == 0
 *   weight  [2][32] float32 (rounded to bf16 inside, as a bf16 module holds it), bias [2] float32
 *   out     [n_images][2][hw] float32: plane 0 = seg map, plane 1 = radius map (the inputs of rcv_vote_frames) */
int rcv_head_1x1(rcv_ctx* ctx, const void* up_bf16, const float* weight, const float* bias, float* out, int n_images, long long hw,
                 void* stream);

/* ---- fused producer head + voting  -- models/fcnresnet.py:118,187-189 + AccumulatorSpace.py:603-656 ----
 * conv8 of the n_kpts radius-map networks -> the evaluator's mask rule -> back-projection -> Accumulator_3D, for n_frames frames:
 * rcv_head_1x1 followed by rcv_vote_frames (sem = the head's seg plane), except that the head's epilogue applies the mask rule
 * itself: the seg plane never reaches HBM, the radius plane is written once and only gathered at the surviving pixels, and
 * no kernel streams the maps a second time.  Results are bit-identical to the two-call sequence.
 *   up     [n_frames][n_kpts][32][H*W] bfloat16: conv7 + BN + ReLU output of keypoint network k for frame f
 *   weight [n_kpts][2][32] float32, bias [n_kpts][2] float32: conv8 of each network (n_kpts <= 8)
 *   depth, K, max_radii, fp, vp and the outputs as in rcv_vote_frames (fp->mask_flags usually RCV_MASK_MAX_RADIUS | RCV_MASK_SEM_GT)
 *   radius_out: optional [n_frames][n_kpts][H*W] float32 to receive the radius planes (NULL: context scratch) */
int rcv_head_vote_frames(rcv_ctx* ctx, int n_frames, int n_kpts, const void* up_bf16, const float* weight, const float* bias,
                         const void* depth, const double* K, const double* max_radii, const rcv_frame_params* fp,
                         const rcv_vote_params* vp, double* centre_mm, int* peak, long long* votes, int* n_points, int* grid, int* status,
                         float* radius_out, void* stream);

/* ---- the producer's tail in one kernel  -- models/fcnresnet.py:114-118 (definition), :183-189 (use) -----------
 * conv7 = Conv2d(64 -> 32, 3x3, padding 1) + BatchNorm2d(32) (eval) + ReLU, then conv8 = Conv2d(32 -> 2, 1x1): an implicit-GEMM
 * tcgen05 kernel (bf16 operands, fp32 accumulation in tensor memory; BatchNorm folded into scale / shift, ReLU, bf16 rounding of
 * the activation and conv8 in the epilogue), so that the 32-channel activation never reaches HBM.
 *   x        [n_images][H][W][64] bfloat16: the output of up1 in channels_last (NHWC) memory, 16-byte aligned; W % 128 == 0
 *   w7       [32][64][3][3] float32 (rounded to bf16 by the kernel)
 *   bn_scale [32] = gamma / sqrt(running_var + eps);  bn_shift [32] = (conv7.bias - running_mean) * bn_scale + beta
 *   w8 [2][32], b8 [2] float32: conv8 (w8 rounded to bf16 like rcv_head_1x1 does)
 *   out      [n_images][2][H*W] float32: plane 0 = seg_pred, plane 1 = radial_pred */
int rcv_conv7_head(rcv_ctx* ctx, const void* x_nhwc_bf16, const float* w7, const float* bn_scale, const float* bn_shift, const float* w8,
                   const float* b8, float* out, int n_images, int height, int width, void* stream);

/* rcv_head_vote_frames with the tail above instead of conv8 alone: x is a HOST array of n_kpts device pointers (the up1 output of
 * each keypoint network, [n_frames][H][W][64] bfloat16 NHWC); w7 [n_kpts][32][64][3][3], bn_scale / bn_shift [n_kpts][32],
 * w8 [n_kpts][2][32], b8 [n_kpts][2].  Everything else as in rcv_head_vote_frames. */
int rcv_conv7_head_vote_frames(rcv_ctx* ctx, int n_frames, int n_kpts, const void* const* x_nhwc_bf16, const float* w7, const float* bn_scale,
                               const float* bn_shift, const float* w8, const float* b8, const void* depth, const double* K,
                               const double* max_radii, const rcv_frame_params* fp, const rcv_vote_params* vp, double* centre_mm, int* peak,
                               long long* votes, int* n_points, int* grid, int* status, float* radius_out, void* stream);

/* ---- introspection ---------------------------------------------------------------------------- */
/* Kernel launches issued by this context since creation (the bench's gpu_launches claim). */
long long rcv_launch_count(const rcv_ctx* ctx);
/* bytes the last rcv_vote_frames_host call copied host -> device (images are cropped to their non-zero depth rows) */
long long rcv_last_h2d_bytes(const rcv_ctx* ctx);
/* Device time of the vote kernel in the most recent rcv_vote_* call, in ms (CUDA events on `stream`);
 * blocks until that call has finished.  Returns a negative value if no call was made. */
float rcv_last_vote_kernel_ms(rcv_ctx* ctx);
/* Device times (ms) of the vote kernel in the most recent n (<= 64) rcv_vote_* calls, oldest first; the
 * events were recorded on the stream each kernel was launched on.  Blocks until they have completed.
 * Returns the number of entries written, or a negative RCV_E_* code. */
int rcv_vote_kernel_times(rcv_ctx* ctx, float* ms_out, int n);
/* Measures this GPU's conflict-free shared-memory atomic rate (atomic lanes per second over all SMs) with a
 * built-in micro-benchmark on the default stream: the denominator of the vote kernel's roofline. */
int rcv_ubench_smem_atomics(rcv_ctx* ctx, double* atomics_per_second);

#ifdef __cplusplus
}
#endif
#endif /* RCVVOTE_H_ */
