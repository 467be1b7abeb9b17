/*
 * cosyb200 - C-ABI of the B200 render-and-compare refinement + multiview matching engine.
 *
 * The reference (ylabbe/cosypose) has no FFI registry for this path: its boundary is the
 * Python class API (cosypose/integrated/pose_predictor.py, multiview_predictor.py) plus one
 * pybind11 module (cosypose/csrc/cosypose_cext.cpp:264-269).  Every entry point below names the
 * reference interface it replaces.  Conventions:
 *   - plain pointers and sizes only; no torch / pybind types cross this boundary;
 *   - every function returns 0 on success and a negative COSYB200_E* code on failure;
 *     cosyb200_last_error() returns a thread-local message for the last failure;
 *   - pointers named *_dev are device pointers on the handle's device, owned by the caller;
 *     pointers named *_host are host pointers; all floating point is IEEE fp32, row-major;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Kernels are only
 *     enqueued; nothing synchronises unless the output is a host array;
 *   - a handle is bound to one device, owns packed weights / mesh tables / workspaces and is
 *     not thread-safe; distinct handles may be used concurrently.
 */
#ifndef COSYB200_H_
#define COSYB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COSYB200_OK 0
#define COSYB200_EINVAL (-1)   /* bad argument (the Python shim raises AssertionError / ValueError) */
#define COSYB200_ECUDA (-2)    /* CUDA runtime error */
#define COSYB200_ESTATE (-3)   /* model / meshes not loaded */
#define COSYB200_ENOMEM (-4)

#define COSYB200_SLOT_COARSE 0
#define COSYB200_SLOT_REFINER 1

#define COSYB200_RENDER_H 240
#define COSYB200_RENDER_W 320
#define COSYB200_N_SAMPLE_POINTS 2000

typedef struct cosyb200_handle cosyb200_handle;

const char* cosyb200_last_error(void);
int cosyb200_version(void);

/* Engine lifetime.  `max_batch` bounds the number of hypotheses per call (the reference chunks
 * at bsz_objects=64, pose_predictor.py:18,34); workspaces are allocated once here. */
int cosyb200_create(cosyb200_handle** out, int device, int max_batch);
int cosyb200_destroy(cosyb200_handle* h);

/* EfficientNet-B3 block table compiled into the engine: out[11] = k,s,e,cin,cexp,cse,cout,
 * pad_lo,pad_hi,skip,n_blocks (reference: models/efficientnet_utils.py:59-81,123-146,259-264). */
int cosyb200_effnet_block(int idx, int32_t* out11);

/* Launch configuration the engine uses for MBConv block `idx` at `batch` hypotheses (host code only, no GPU
 * needed; tests/test_host_logic.py checks the resource budgets): out[32] =
 *   [0..7]   tiled depthwise: used, rows per tile, row strips, x tiles, unit width, units per row, channel chunks,
 *            dynamic shared memory bytes;  [8..11] rolling depthwise: tiles, channel chunks, threads, rows per tile;
 *   [12..19] expand 1x1 (tensor cores): n tile width, n tiles, k stages, weight slots, resident, dynamic shared
 *            memory bytes, producer groups, output tiles;  [20..27] project 1x1: the same. */
int cosyb200_launch_plan(int idx, int batch, int32_t* out32);

/* Work split of the 3xFP16 1x1-convolution kernel (gemm_impl = 2, the default) for C[M][N] = A[M][K] W^T on a GPU with
 * n_sms SMs (host code only): out[9] = n tile width, small-tile kernel shape, n tiles, k stages (of 32), weight slots,
 * weights resident, dynamic shared memory bytes, grid (CTAs), split_k.  split_k = 1: the CTAs of an n-tile column take
 * equal contiguous ranges of its (m-tile, k-stage) units, so an m-tile may be shared by two consecutive CTAs (fixed-order
 * partial-sum hand-over); 0: whole m-tiles per CTA.  tests/test_host_logic.py checks the invariants. */
int cosyb200_pw2_plan(int M, int N, int K, int n_sms, int32_t* out9);

/* Replaces PosePredictor.load_state_dict (reference: models/pose.py:18-36, weights named as in
 * SURVEY.md section 5).  `names[i]` is a state_dict key, `ptrs_host[i]` its fp32 data, `numels[i]`
 * its element count.  BatchNorm (eps 1e-3) is folded into the conv weights here. */
int cosyb200_load_pose_model(cosyb200_handle* h, int slot, int n_tensors,
                             const char* const* names, const float* const* ptrs_host,
                             const int64_t* numels);

/* Replaces BatchedMeshes (reference: lib3d/rigid_mesh_database.py:59-79) + the deterministic
 * 2000-point subset of Meshes.sample_points (lib3d/mesh_ops.py:31-41).
 *   points_host [n_labels, n_points, 3]; point_ids_host [n_sample] (the RandomState(0) subset);
 *   sym_host [n_labels, s_max, 4, 4] identity padded; n_sym_host [n_labels];
 *   aabb_host [n_labels, 8, 3] (lib3d/mesh_ops.py:15-28) - the points RANSAC scoring uses. */
int cosyb200_set_meshes(cosyb200_handle* h, int n_labels, int n_points, const float* points_host,
                        int n_sample, const int64_t* point_ids_host, int s_max,
                        const float* sym_host, const int32_t* n_sym_host, const float* aabb_host);

/* TCO_init_from_boxes(z_range=(1,1)) (reference: lib3d/cosypose_ops.py:121-135) when zup == 0,
 * TCO_init_from_boxes_zup_autodepth (cosypose_ops.py:138-173) when zup == 1.
 *   boxes_dev [B,4], K_dev [B,3,3] (per hypothesis), label_ids_dev [B] -> TCO_dev [B,4,4]. */
int cosyb200_tco_init(cosyb200_handle* h, int B, int zup, const float* boxes_dev,
                      const float* K_dev, const int32_t* label_ids_dev, float* TCO_dev,
                      void* stream);

/* Phase A of one iteration = PosePredictor.crop_inputs without the pixel crop
 * (reference: models/pose.py:45-67 -> lib3d/camera_geometry.py:18-87, lib3d/cropping.py:7-47).
 *   K_dev [B,3,3] per hypothesis, TCO_dev [B,4,4], label_ids_dev [B]
 *   -> boxes_rend_dev [B,4], boxes_crop_dev [B,4], K_crop_dev [B,3,3].
 * The host may then call renderer.render(obj_infos, TCO, K_crop) exactly as pose.py:100-102. */
int cosyb200_prepare_iter(cosyb200_handle* h, int B, int img_h, int img_w, const float* K_dev,
                          const float* TCO_dev, const int32_t* label_ids_dev,
                          float* boxes_rend_dev, float* boxes_crop_dev, float* K_crop_dev,
                          void* stream);

/* The RoI crop alone (reference: lib3d/cropping.py:74, torchvision.ops.roi_align with
 * output (240,320), spatial_scale 1, sampling_ratio 4, aligned False).
 *   images_dev [n_images,3,img_h,img_w] NCHW, im_ids_dev [B], boxes_crop_dev [B,4]
 *   -> crops_dev [B,3,240,320] NCHW. */
int cosyb200_roi_crop(cosyb200_handle* h, int B, const float* images_dev, int n_images, int img_h,
                      int img_w, const int32_t* im_ids_dev, const float* boxes_crop_dev,
                      float* crops_dev, void* stream);

/* PosePredictor.net_forward on an explicit 6-channel input (reference: models/pose.py:81-87,
 * models/efficientnet.py:174-190).  crops_dev / renders_dev [B,3,240,320] NCHW are channels
 * 0-2 / 3-5 of the concatenated input (pose.py:104) -> pose9_dev [B,9].
 * `taps_dev` (may be NULL) is an array of 28 device pointers (stem, block0..25, head) that
 * receive the block-boundary activations in NHWC; NULL entries are skipped. */
int cosyb200_net_forward(cosyb200_handle* h, int slot, int B, const float* crops_dev,
                         const float* renders_dev, float* pose9_dev, float* const* taps_dev,
                         void* stream);

/* PosePredictor.update_pose, pose_dim 9 (reference: models/pose.py:69-79, lib3d/rotations.py:6-21,
 * lib3d/cosypose_ops.py:10-31).  TCO_in [B,4,4], K_crop [B,3,3], pose9 [B,9] -> TCO_out [B,4,4]. */
int cosyb200_update_pose(cosyb200_handle* h, int B, const float* TCO_in_dev,
                         const float* K_crop_dev, const float* pose9_dev, float* TCO_out_dev,
                         void* stream);

/* Phase B of one iteration: crop + concat + backbone + head + pose update
 * (reference: models/pose.py:99-108 minus the renderer call).  renders_dev is fp32 NCHW
 * [B,3,240,320] in [0,1] (render_u8 == 0, the tensor pose.py:100 receives) or uint8 NHWC
 * [B,240,320,3] (render_u8 == 1, what the renderer produces before `.float()/255`,
 * rendering/bullet_batch_renderer.py:70-83; converted inside the stem kernel).  Images are gathered through
 * im_ids_dev instead of being copied per hypothesis (pose_predictor.py:41). */
int cosyb200_refine_iter(cosyb200_handle* h, int slot, int B, const float* images_dev,
                         int n_images, int img_h, int img_w, const int32_t* im_ids_dev,
                         const float* boxes_crop_dev, const void* renders_dev, int render_u8,
                         const float* K_crop_dev, const float* TCO_in_dev, float* pose9_dev,
                         float* TCO_out_dev, void* stream);

/* ---- device rasteriser ----
 * Replaces the renderer side input of the loop (reference: models/pose.py:100-102 -> rendering/bullet_batch_renderer.py:46-90:
 * one pybullet getCameraImage per hypothesis in worker processes, frames back through a multiprocessing queue and a
 * pinned copy).  Camera as simulator/camera.py:10-34 builds it from K: pinhole, samples at pixel centres (j+0.5, i+0.5),
 * near plane 0.01 (triangles reaching in front of it are dropped, not clipped), nearest surface wins, background 0,
 * both faces drawn; colour = perspective-correct interpolation of per-vertex colours, unlit (pybullet's shading is not
 * reproducible here: DESIGN.md).
 *   cosyb200_set_render_meshes   host tables: vertices / colors [n_vertices,3] (object frame, metres; colours in [0,1]),
 *       faces [n_faces,3] int32 ids into the vertex table, face_offsets [n_labels+1]: label l owns faces
 *       face_offsets[l] .. face_offsets[l+1]-1 (label ids as in cosyb200_set_meshes).
 *   cosyb200_render   B views at TCO_dev [B,4,4] with intrinsics K_dev [B,3,3] (K_crop of cosyb200_prepare_iter):
 *       out_u8 = 1 -> uint8 [B,240,320,3] (the layout refine_iter takes with render_u8 = 1),
 *       out_u8 = 0 -> fp32 [B,3,240,320] = uint8 / 255 as bullet_batch_renderer.py:83 returns it;
 *       depth_dev (may be NULL) [B,240,320]: camera-frame z of the visible surface in metres, 0 on the background
 *       (render_depth=True of the reference, bullet_scene_renderer.py:51-56). */
int cosyb200_set_render_meshes(cosyb200_handle* h, int n_labels, int64_t n_vertices, const float* vertices,
                               const float* colors, int64_t n_faces, const int32_t* faces,
                               const int32_t* face_offsets);
int cosyb200_render(cosyb200_handle* h, int B, const int32_t* label_ids_dev, const float* TCO_dev,
                    const float* K_dev, void* out_dev, int out_u8, float* depth_dev, void* stream);

/* PosePredictor.forward with pre-rendered views (reference: models/pose.py:89-132): n_iter
 * iterations without returning to the host.  renders_dev [n_iter,B,3,240,320]; K_dev [B,3,3];
 * outputs are per iteration: TCO_out [n_iter,B,4,4], K_crop [n_iter,B,3,3], boxes_rend and
 * boxes_crop [n_iter,B,4], pose9 [n_iter,B,9]; iteration n reads TCO_out[n-1] (TCO_in_dev for n=0).
 * renders_dev == NULL: every iteration rasterises its own views on the device from the meshes of
 * cosyb200_set_render_meshes at the iteration's input poses and K_crop (the reference's loop, pose.py:99-102,
 * with the renderer inside the engine: nothing returns to the host between iterations). */
int cosyb200_refine_n(cosyb200_handle* h, int slot, int B, int n_iter, const float* images_dev,
                      int n_images, int img_h, int img_w, const int32_t* im_ids_dev,
                      const float* K_dev, const int32_t* label_ids_dev, const void* renders_dev,
                      int render_u8, const float* TCO_in_dev, float* TCO_out_dev, float* K_crop_dev,
                      float* boxes_rend_dev, float* boxes_crop_dev, float* pose9_dev, void* stream);

/* Engine options (all choose between implementations of the same arithmetic; results agree within the
 * tolerances of tests/test_gpu_*.py):
 *   "gemm_impl": 2 (default) runs the 1x1 convolutions on the tcgen05 tensor cores with the 3xFP16 hi/lo split
 *                (kernels_pw2.cuh), 1 = the 3xTF32 kernel, 0 = CUDA cores in plain fp32 (the per-block parity anchor);
 *   "xdw":       1 (default) blocks 2-8 run expand 1x1 + depthwise + pooling as one kernel (kernels_xdw.cuh), 0 = separately;
 *   "dw_impl":   1 (default) shared-memory tiled depthwise + split squeeze-excite for the blocks with
 *                output <= 30x40, 0 = rolling-window depthwise + k_se_gate everywhere;
 *   "tc_groups": 0 (default) picks the 3xTF32 kernel variant per layer, 1 / 2 force one / two producer warpgroups;
 *   "graph":     1 (default) cosyb200_refine_n replays a captured CUDA graph when its arguments repeat, 0 = plain launches;
 *   "trace_block": debugging, -1 (default) off: the fused kernel of that block stamps its phases (cosyb200_debug_trace). */
int cosyb200_set_option(cosyb200_handle* h, const char* name, int value);

/* One 1x1 convolution on caller data, for kernel-level tests (reference: the Conv2d 1x1 + folded
 * BatchNorm (+ swish) of models/efficientnet.py:80-81,90,188):
 *   C[M,N] = act((A[M,K] * gate[m / rows_per_img, k]) @ W[N,K]^T + bias[N]) (+ resid[M,N])
 * A, gate (may be NULL), resid (may be NULL), C on the device; W, bias on the host.  Synchronises. */
int cosyb200_debug_pointwise(cosyb200_handle* h, int impl, int M, int N, int K, const float* A_dev,
                             const float* W_nk_host, const float* bias_host, const float* gate_dev,
                             int rows_per_img, const float* resid_dev, int swish, float* C_dev,
                             void* stream);

/* Tuning aid: when trace_dev (32 int64 slots, device memory) is non-NULL, CTA (0,0) of the tensor-core
 * 1x1 kernel stores clock64() stamps of its pipeline events there; NULL switches it off. */
int cosyb200_debug_trace(cosyb200_handle* h, long long* trace_dev);
/* Debugging aid: during the following trunk forwards, copy MBConv block `block`'s expanded activation
   [B*Hin*Win][Cexp], depthwise output [B*Hout*Wout][Cexp] and squeeze-excite gate [B][Cexp] into the given
   device buffers (NULL = skip; block -1 = off). */
int cosyb200_debug_dump(cosyb200_handle* h, int block, float* expanded, float* dw_out, float* gate);

/* Launch accounting (no reference counterpart; the reference times with a wall-clock Timer,
 * utils/timer.py:4-36).  Every kernel the engine launches is counted per category:
 *   0 geometry, 1 roi crop, 2 stem, 3 expand 1x1, 4 depthwise, 5 squeeze-excite, 6 project 1x1,
 *   7 head 1x1, 8 pool+fc+update, 9 ransac, 10 rasteriser.
 * With profiling enabled each launch is also bracketed by CUDA events on its stream and the
 * device time accumulated per category (adds launch overhead: use outside timed regions). */
int cosyb200_profile_enable(cosyb200_handle* h, int on);
int cosyb200_profile_read(cosyb200_handle* h, int reset, int64_t* launches11, double* ms11);
/* The same device time split per MBConv block: ms[category * 32 + block], block 31 = stem / head / outside. */
int cosyb200_profile_read_blocks(cosyb200_handle* h, int reset, double* ms352);

/* ---- multiview candidate matching (reference: multiview/ransac.py:137-199) ---- */

/* cosypose_cext.make_ransac_infos (reference: csrc/cosypose_cext.cpp:36-105).  Labels are dense
 * ids (label equality is all the reference uses).  Two calls: with seeds_host == NULL it only
 * counts; then the caller allocates and calls again.
 *   seeds_host [6, n_seeds] rows: view1, view2, match1_cand1, match1_cand2, match2_cand1,
 *   match2_cand2;  tmatches_host [3, n_tmatches] rows: hypothesis_id, cand1, cand2. */
int cosyb200_ransac_infos(int n_cand, const int32_t* view_ids_host, const int32_t* label_ids_host,
                          int n_ransac_iter, int seed, int64_t* n_seeds, int64_t* n_tmatches,
                          int32_t* seeds_host, int32_t* tmatches_host);

/* estimate_camera_poses over all seeds (reference: multiview/ransac.py:19-64).
 *   poses_dev [n_cand,4,4], cand_label_ids_dev [n_cand], seeds_dev [6,n_seeds] (layout above)
 *   -> TC1C2_dev [n_seeds,4,4]. */
int cosyb200_ransac_models(cosyb200_handle* h, int64_t n_seeds, const float* poses_dev,
                           const int32_t* cand_label_ids_dev, const int32_t* seeds_dev,
                           float* TC1C2_dev, void* stream);

/* score_tmatches over all rows (reference: multiview/ransac.py:67-88).
 *   tmatches_dev [3,n_tmatches], TC1C2_dev [n_seeds,4,4] -> dists_dev [n_tmatches]. */
int cosyb200_ransac_score(cosyb200_handle* h, int64_t n_tmatches, const float* poses_dev,
                          const int32_t* cand_label_ids_dev, const int32_t* tmatches_dev,
                          const float* TC1C2_dev, float* dists_dev, void* stream);

/* symmetric_distance_batched_fast (reference: lib3d/symmetric_distances.py:38-57) on the AABB
 * points: T1_dev, T2_dev [n,4,4], label_ids_dev [n] -> dists_dev [n], best_sym_dev [n] (may be NULL). */
int cosyb200_symmetric_distance(cosyb200_handle* h, int64_t n, const float* T1_dev,
                                const float* T2_dev, const int32_t* label_ids_dev,
                                float* dists_dev, int32_t* best_sym_dev, void* stream);

/* cosypose_cext.find_ransac_inliers (reference: csrc/cosypose_cext.cpp:107-216).  Outputs are
 * written into caller buffers of capacity n_tmatches (matches) / n_seeds (best hypotheses);
 * the counts come back through n_inlier_matches / n_best. */
int cosyb200_ransac_inliers(int64_t n_seeds, const int32_t* seeds_view1_host,
                            const int32_t* seeds_view2_host, int64_t n_tmatches,
                            const int32_t* mtc_hypothesis_id_host, const int32_t* mtc_cand1_host,
                            const int32_t* mtc_cand2_host, const float* dists_host,
                            float dist_threshold, int n_min_inliers, int32_t* inlier_cand1_host,
                            int32_t* inlier_cand2_host, int64_t* n_inlier_matches,
                            int32_t* best_hypotheses_host, int64_t* n_best);

/* The same voting on the device (no device -> host copy of the distances; reference: csrc/cosypose_cext.cpp:107-216,
 * including the stable tie order and the id-0 quirk at :203).  Seeds must be grouped by ordered view pair in ascending
 * (view1, view2) order, as cosyb200_ransac_infos emits them: pair_start_dev [n_pairs + 1] are the seed offsets of the
 * pairs; the rows of one hypothesis must be contiguous and hypothesis ids ascending (also as emitted).
 * out_c1 / out_c2 [capacity n_mtc], best [capacity n_pairs], counts [2] = {#inlier matches, #kept pairs} (int64). */
int cosyb200_ransac_inliers_dev(cosyb200_handle* h, int64_t n_seeds, int64_t n_pairs,
                                const int32_t* pair_start_dev, int64_t n_tmatches,
                                const int32_t* tmatches_hyp_dev, const int32_t* tmatches_cand1_dev,
                                const int32_t* tmatches_cand2_dev, const float* dists_dev,
                                float dist_threshold, int n_min_inliers, int32_t* inlier_cand1_dev,
                                int32_t* inlier_cand2_dev, int32_t* best_hypotheses_dev, int64_t* counts_dev,
                                void* stream);

/* cosypose_cext.scatter_argmin (reference: csrc/cosypose_cext.cpp:218-245): first minimum per
 * group; out_host has one entry per group id 0..n_groups-1. */
int cosyb200_scatter_argmin(int64_t n, const float* values_host, const int32_t* group_ids_host,
                            int64_t n_groups, int32_t* out_host);

/* cosypose_cext.expand_ids_for_symmetry (reference: csrc/cosypose_cext.cpp:247-259) over dense
 * label ids; returns the expanded length through n_out (call with ids_expand_host == NULL to count). */
int cosyb200_expand_ids_for_symmetry(int64_t n, const int32_t* label_ids_host,
                                     const int32_t* n_sym_per_label_host, int64_t* n_out,
                                     int32_t* ids_expand_host, int32_t* sym_ids_host);

/* out[i] = inv(A[ia[i]]) @ B[ib[i]] for rigid 4x4 transforms (ia / ib may be NULL = identity map):
 * `invert_T(TWC) @ TWO` of reproject_scene (reference: integrated/multiview_predictor.py:20-41) and
 * the known-camera branch `invert_T(TWC1) @ TWC2` (multiview/ransac.py:169-173). */
int cosyb200_compose_inv(cosyb200_handle* h, int64_t n, const float* A_dev, const int32_t* ia_dev,
                         const float* B_dev, const int32_t* ib_dev, float* out_dev, void* stream);

/* One linearisation of the object-level bundle adjustment problem = MultiviewRefinement.
 * align_TCO_cand + forward_jacobian (reference: multiview/bundle_adjustment.py:164-214) with analytic
 * derivatives instead of autograd, plus the normal equations of compute_lm_step (:216-222).
 *   cand_TCO [n_cand,4,4], cand_obj / cand_view / cand_label [n_cand] (local object / view ids,
 *   dense label ids), TWO_9d [n_obj,9], TCW_9d [n_view,9], K [n_view,3,3], points [n_labels,n_pts,3]
 *   -> align_dists [n_cand], aligned [n_cand,4,4] (candidate poses times their best symmetry),
 *      errors [n_res] (n_res = n_cand*n_pts*2, order candidate, point, x|y), Jc [n_res,18]
 *      (d yhat / d TWO_9d[obj] | d yhat / d TCW_9d[view]), loss [1] = mean(min(e^2, threshold)),
 *      and when non-NULL JtJ [n_params,n_params], Jte [n_params] with n_params = 9*(n_obj+n_view),
 *      object parameters first.  The small dense solve stays with the caller, as in the reference. */
int cosyb200_ba_linearize(cosyb200_handle* h, int n_cand, int n_obj, int n_view, int n_pts,
                          const float* cand_TCO_dev, const int32_t* cand_obj_dev,
                          const int32_t* cand_view_dev, const int32_t* cand_label_dev,
                          const float* TWO_9d_dev, const float* TCW_9d_dev, const float* K_dev,
                          const float* points_dev, float residuals_threshold, float* align_dists_dev,
                          float* aligned_dev, float* errors_dev, float* Jc_dev, float* JtJ_dev,
                          float* Jte_dev, float* loss_dev, void* stream);

/* The same linearisation evaluated in float64 on the device (residuals, Jacobian, normal equations, loss): the LM
 * normal equations are ill conditioned in fp32 - the reference's own fp32 result is ~1e-4 from the same code run in
 * float64 - and the problem is tiny.  Parameters and the aligned poses stay fp32; JtJ64 [n_params^2], Jte64
 * [n_params] (both may be NULL), loss64 [1] are float64 device buffers of the caller. */
int cosyb200_ba_linearize_f64(cosyb200_handle* h, int n_cand, int n_obj, int n_view, int n_pts,
                              const float* cand_TCO_dev, const int32_t* cand_obj_dev,
                              const int32_t* cand_view_dev, const int32_t* cand_label_dev,
                              const float* TWO_9d_dev, const float* TCW_9d_dev, const float* K_dev,
                              const float* points_dev, float residuals_threshold, float* align_dists_dev,
                              float* aligned_dev, double* JtJ64_dev, double* Jte64_dev, double* loss64_dev,
                              void* stream);

/* compute_lm_step (reference: multiview/bundle_adjustment.py:216-222, `pinverse(JtJ + lambda I) @ Jte` on the CPU):
 * float64 Cholesky solve on the device, one CTA; step_dev [n] fp32.  n_bad_pivots_dev (may be NULL) receives the
 * number of non-positive pivots that had to be replaced (0 for a positive definite system). */
int cosyb200_lm_solve(cosyb200_handle* h, int n, const double* JtJ64_dev, const double* Jte64_dev, double lambda,
                      float* step_dev, int32_t* n_bad_pivots_dev, void* stream);

/* ADD / ADD-S pose errors (reference: lib3d/distances.py:5-21 and the statistics of
 * evaluation/meters/pose_meters.py:84-89).  T_pred, T_gt [n,4,4], points [n,P,3] (the label's model points per
 * pair), symmetric [n] (0: ADD, else ADD-S: every ground-truth point against its closest predicted point, first minimum, as `dists_add_symmetric`; NULL = all ADD)
 * -> dists [n,P,3] (may be NULL), norm_avg [n], xyz_avg [n,3], TCO_xyz [n,3], TCO_norm [n]. */
int cosyb200_pose_errors(cosyb200_handle* h, int n, int n_points, const float* T_pred_dev,
                         const float* T_gt_dev, const float* points_dev, const int32_t* symmetric_dev,
                         float* dists_dev, float* norm_avg_dev, float* xyz_avg_dev, float* tco_xyz_dev,
                         float* tco_norm_dev, void* stream);

/* Index preconditions.  label ids and image ids of the single-view entry points (tco_init, prepare_iter, roi_crop,
 * refine_iter, refine_n) are clamped into their tables on the device, so a bad id cannot read out of bounds (its result
 * is meaningless).  The multiview entry points (ransac_models / ransac_score / symmetric_distance / ransac_inliers_dev /
 * ba_linearize*) take candidate, label, object and view ids that MUST be valid: they index caller-sized arrays whose
 * lengths the ABI does not carry; the Python shim builds them from cosyb200_ransac_infos and dense label tables. */

/* ---- multi-GPU exchange ----------------------------------------------------------
 * Hypotheses shard across ranks with no data-path collective; the ONE exchange is an all-gather of fixed-size fp32
 * records per hypothesis after the last refinement iteration.  The reference gathers per-rank predictions through
 * pickle files on a shared filesystem (utils/tensor_collection.py:142-163, datasets/samplers.py:20-34).
 * NCCL is bound at run time (dlopen of the libnccl already in the process, e.g. PyTorch's), so the library loads
 * and every other entry point works on hosts without NCCL.
 *   cosyb200_nccl_unique_id   rank 0 creates the 128-byte id; the caller broadcasts it (any transport)
 *   cosyb200_nccl_comm_init   every rank: communicator of `world` ranks for the handle's device
 *   cosyb200_allgather_candidates   all_dev [world * count_per_rank] floats; rank r's records live at
 *       all_dev + r * count_per_rank.  local_dev == NULL means "already there" (in-place: pass the shard's slice of
 *       all_dev as the output buffer of cosyb200_refine_n and nothing is staged); enqueued on `stream`. */
int cosyb200_nccl_unique_id(char* id128_host);
int cosyb200_nccl_comm_init(cosyb200_handle* h, int world, int rank, const char* id128_host);
int cosyb200_nccl_comm_destroy(cosyb200_handle* h);
int cosyb200_allgather_candidates(cosyb200_handle* h, const float* local_dev, float* all_dev,
                                  int64_t count_per_rank, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COSYB200_H_ */
