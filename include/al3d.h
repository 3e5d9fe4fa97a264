/*
 * al3d.h -- C ABI of libal3d.so, the B200 (sm_100a) implementation of the 3DAL object-centric
 * auto-labeling hot path (Frustum-PointNet static / dynamic box refinement + points-in-box crop).
 *
 * Conventions:
 *   - every function returns 0 on success, non-zero on failure; al3d_last_error() returns the text
 *     of the last failure on the calling thread.  Nothing throws across this boundary.
 *   - all pointers are DEVICE pointers unless the name ends in _host.  The library allocates
 *     nothing: inputs, outputs, packed weights and scratch are owned by the caller (PyTorch).
 *   - everything is stream-ordered on `stream` (a cudaStream_t passed as void*); there are no
 *     hidden synchronisations.  Process state is limited to: the per-thread error string, one 64-byte
 *     pinned watchdog status block per device (al3d_tc_abort_code) and the two diagnostic switches
 *     al3d_tc_configure / al3d_set_debug_buffer.
 *   - tensors are dense row-major unless strides are passed (in ELEMENTS).
 *
 * Each entry point names the reference code (jacky121298/3DAL_PyTorch, file:line) it replaces.
 * The reference-side binding (ctypes) is shown in INTEGRATION.md and implemented in
 * 3dal_pytorch_b200/_lib.py.
 */
#ifndef AL3D_H_
#define AL3D_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AL3D_ABI_VERSION 3

/* activation flags for al3d_linear_f32 */
#define AL3D_ACT_NONE 0
#define AL3D_ACT_RELU 1
#define AL3D_ACT_ACCUMULATE 2   /* y += result (training backward: a tensor with two consumers) */

/* gather index policies (al3d_gather_fg) */
#define AL3D_GATHER_STRIDED 0   /* device rule: slot j <- pos[(j*L)/n_pts] (L>=n_pts) or pos[j%L]   */
#define AL3D_GATHER_TABLE   1   /* caller-provided (bs,n_pts) int32 choice table (numpy RNG replay) */

int         al3d_abi_version(void);
const char *al3d_last_error(void);
/* 1 if the current device is sm_100 (B200) and the tcgen05 kernels can run, else 0. */
int         al3d_device_supports_tcgen05(void);

/* ------------------------------------------------------------------------------------------------
 * Shared-MLP building blocks, fp32 SIMT ("exact" precision mode).
 * Replaces nn.Conv1d(k=1)/nn.Linear + eval BatchNorm1d (folded by the caller) + ReLU
 * (tools/static_model.py:279-283,289-294,330-338; tools/dynamic_model.py:241-248,278-285,307-311).
 * ---------------------------------------------------------------------------------------------- */

/* First layer on a strided (bs,C,n) point tensor: y[(b*n+p), o] = act(sum_c x[b,c,p]*w[o,c] + bias[o]).
 * x strides in elements; C <= 8.  y is (bs*n, cout) row-major. */
int al3d_pointwise_first_f32(const float *x, int64_t sb, int64_t sc, int64_t sp, int bs, int C, int n,
                             const float *w, const float *bias, int cout, int act, float *y, void *stream);

/* y[m, o] = act(sum_k a[m*lda + k] * w[o*ldw + k] + bias[o] + rowbias[(m / rows_per_group)*cout + o]).
 * bias and rowbias may be NULL.  If y_max != NULL the (M,cout) result is NOT stored; instead
 * y_max[(m / rows_per_group)*cout + o] = max(existing, result) is accumulated with an integer
 * atomicMax, which requires act == RELU and y_max pre-filled with zeros (max-pool over the points
 * of an object, tools/static_model.py:284,334). */
int al3d_linear_f32(const float *a, int64_t lda, int64_t M, int K, const float *w, int64_t ldw,
                    const float *bias, const float *rowbias, int64_t rows_per_group, int cout, int act,
                    float *y, int64_t ldy, float *y_max, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Foreground mask, ordered compaction and gather (integer / indexing work, bit-exact).
 * Replaces point_cloud_masking + gather_object_pts (tools/static_model.py:23-62,
 * tools/dynamic_model.py:24-63).
 * ---------------------------------------------------------------------------------------------- */

/* mask[b,p] = logits[b,p,0] < logits[b,p,1] (strict, NaN -> 0) when logits != NULL, otherwise
 * mask is read as given.  pos[b, 0:count[b]] = ascending indices p with mask set. */
int al3d_mask_compact(const float *logits, uint8_t *mask, int bs, int n, int32_t *pos, int32_t *count, void *stream);

/* out[b,c,j] = x[b,c,pos[b,sel(j)]] for j < n_pts (zeros when count[b]==0); indices[b,j] likewise
 * (may be NULL).  policy AL3D_GATHER_STRIDED computes sel on the device, AL3D_GATHER_TABLE reads
 * choice[b,j].  x strides in elements. */
int al3d_gather_fg(const float *x, int64_t sb, int64_t sc, int64_t sp, int bs, int C, int n,
                   const int32_t *pos, const int32_t *count, int policy, const int32_t *choice, int n_pts,
                   float *out, int64_t *indices, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Head parsing, decode and the two-stage canonical re-transform.
 * ---------------------------------------------------------------------------------------------- */

/* parse_output_to_tensors (tools/static_model.py:64-96) + the centre residual add (:132,174,211).
 * box_pred (bs,39).  add (bs, add_stride) may be NULL; centre_out = box_pred[:, :3] + add[:, :3].
 * Any output pointer may be NULL. */
int al3d_parse_heads(const float *box_pred, int bs, const float *add, int64_t add_stride,
                     float *center_boxnet, float *center, float *heading_scores, float *heading_res_norm,
                     float *heading_res, float *size_scores, float *size_res_norm, float *size_res, void *stream);

/* Box decode shared by the two-box forward (tools/static_model.py:178-190) and the eval loops
 * (tools/static_eval.py:270-288, tools/dynamic_eval.py:227-242): argmax of the heading / size
 * scores, class2angle (+wrap) / class2size in float64 (tools/utils.py:69-79), plus base heading.
 * box_out (bs,7) f32 = [center, size, heading]; cls_out (bs,2) int32 = [heading bin, size cluster]
 * (may be NULL). */
int al3d_decode_boxes(const float *center, const float *heading_scores, const float *heading_res,
                      const float *size_scores, const float *size_res, const float *base_heading,
                      int64_t base_stride, int bs, float *box_out, int32_t *cls_out, void *stream);

/* Two-stage re-centering (tools/static_model.py:195-205): P <- Rz(-h1) (Rz(h0) P + c0 - c1) on the
 * gathered (bs,3,m) points, and the head-two heading label angle2class(gt_heading - h1)
 * (tools/utils.py:53-60 evaluated in f32 like the reference's 0-dim tensors). */
int al3d_twostage_retransform(const float *obj_pts, int bs, int m, const float *init_box, const float *box_one,
                              const float *bbox_gt, float *obj_pts_two, int64_t *heading_cls_label,
                              float *heading_res_label, void *stream);


/* Fused FC chain of a box head / embedding in one launch (csrc/heads.cu): up to three Linear (+ folded BatchNorm)
 * (+ ReLU) layers on (bs, k0 [+ k1]) fp32 rows (x1: optional second input, concatenated -- the dynamic head's
 * cat[point_e, box_e], tools/dynamic_model.py:137), weights TRANSPOSED (K, N) row-major.  With heads = 1 the last layer is
 * the 39-wide head vector and the epilogue does parse_output_to_tensors (tools/static_model.py:64-96), the centre
 * residual add (:132,174,211) and, when box != NULL, the decode of al3d_decode_boxes (tools/static_eval.py:270-288).
 * Replaces PointNetEstimation fc1-fc3 (tools/static_model.py:336-338; tools/dynamic_model.py:307-311) and the embedding
 * FC layers (tools/dynamic_model.py:247-248, 284-285).  Every output pointer may be NULL. */
typedef struct al3d_fc_chain_desc {
    const float *x0; int32_t k0; int32_t pad0; int64_t ld0;
    const float *x1; int32_t k1; int32_t pad1; int64_t ld1;
    int32_t n_layers; int32_t width[3]; int32_t relu[3]; int32_t heads;
    const float *wt[3]; const float *bias[3];
    float *out; int64_t ldo;
    const float *add; int64_t add_stride;
    const float *base_heading; int64_t base_stride;
    float *center_boxnet, *center, *heading_scores, *heading_res_norm, *heading_res, *size_scores, *size_res_norm, *size_res, *box;
    int32_t *cls;
} al3d_fc_chain_desc;
int al3d_fc_chain(const al3d_fc_chain_desc *d, int bs, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Points-in-rotated-box crop over a batch of lidar frames (integer / indexing work, bit-exact).
 * Replaces box_np_ops.points_in_rbbox (det3d/core/bbox/box_np_ops.py:641-647 ->
 * det3d/core/bbox/geometry.py:241-276) and the per-box crop loop of _create_pd_detection
 * (det3d/datasets/waymo/waymo_common.py:167-171).  The points of frame f are the (N_f, pt_stride) f32 rows at
 * frame_points[f] (a device array of F device pointers: frames are used where they lie, nothing is concatenated; x y z
 * are the first three columns); the boxes of all frames are concatenated: inward plane equations planes (sum B_f, 6, 4) f32 and padded
 * axis-aligned rectangles aabb (sum B_f, 6) f32 [xmin ymin zmin xmax ymax zmax] with box_off (F+1) i64
 * (both from al3d_crop_box_setup).  Every point ends up in exactly the lists the reference predicate puts it in.
 * `overflow` is a device int32 the kernels set non-zero when a caller-provided capacity is too small.
 * Order of calls: build_grid -> hits -> scan -> (read offsets[n_boxes] = total, allocate) -> fill.
 * ---------------------------------------------------------------------------------------------- */
/* boxes (n,7) f32 [x y z l w h heading] + sincos (n,2) f32 [sin, cos of the heading, from the host's float32
 * numpy] -> planes (n,6,4) f32 and padded rectangles aabb (n,6) f32 (pad = pad_abs + pad_rel * max|corner|), bit-identical
 * to the reference's numpy arithmetic (box_np_ops.py:55-85,146-179,241-262,650-670; geometry.py:351-377). */
int al3d_crop_box_setup(const float *boxes, const float *sincos, int64_t n_boxes, float pad_abs, float pad_rel, float *planes,
                        float *aabb, void *stream);
/* boxes + sincos as above -> local (n,12) f32, 16-byte aligned: the box in its own frame [cx cy cz cos | sin l/2 w/2 h/2 |
 * margin - - -] for the conservative pair classification of al3d_crop_hits (only pairs within the rounding margin of a
 * face evaluate the exact plane predicate; margin = NaN marks a box that always does). */
int al3d_crop_box_local(const float *boxes, const float *sincos, int64_t n_boxes, float *local, void *stream);
int al3d_crop_chunk_points(void);      /* points per work chunk (the caller builds the chunk table) */
int al3d_crop_occ_words(void);         /* 32-bit words of one frame's fine occupancy bitmap */
int al3d_crop_hit_bytes(void);         /* bytes of one hit record of al3d_crop_hits / al3d_crop_fill */
/* Per frame: grid_meta (n_frames, 8) f32, the coarse G x G cell -> box lists (CSR: cell_start (n_frames, G*G+1),
 * cell_boxes (n_frames, cell_cap); with cell_cap = 0 nothing is stored and cell_start[f, G*G] receives the capacity
 * the frame needs), when cell4 != NULL the packed cell entries cell4 (n_frames, G*G, 2) u32 [id0 | id1 << 16,
 * id2 | count << 16; count 0xFFFF = use the CSR list; max_boxes must be the value al3d_crop_hits gets: a build with
 * CROP_C4_SMEM stores (n_frames, G*G) 4-byte words id0 | id1 << 10 | id2 << 20 | count << 30 instead when it is <= 256] and, when occ != NULL, the fine occupancy bitmap occ (n_frames,
 * al3d_crop_occ_words()) u32 rasterised from the rotated box footprints (boxes (n,7) + sincos (n,2) as for
 * al3d_crop_box_setup). */
int al3d_crop_build_grid(const float *aabb, const float *boxes, const float *sincos, const int64_t *box_off, int n_frames, int G,
                         float *grid_meta, int32_t *cell_start, int32_t *cell_boxes, int cell_cap, uint32_t *cell4, int max_boxes,
                         uint32_t *occ, int32_t *overflow, void *stream);
/* chunks: (n_chunks, 4) i32 rows [frame, first point in frame, n points, chunk index in frame];
 * hits: scratch of n_chunks * 8 * hit_cap slots of al3d_crop_hit_bytes() bytes, 16-byte aligned (one ordered segment per
 * warp of the chunk's CTA; stored as an array of 16-byte records x y z | point index followed by an array of 4-byte
 * words box | rank << 16), n_hits (n_chunks, 8) i32;
 * chunk_box_count: (n_chunks, max_boxes) i32 scratch. */
int al3d_crop_hits(const float *const *frame_points, int64_t pt_stride, const float *planes,
                   const float *local, const int64_t *box_off, int G, const float *grid_meta, const int32_t *cell_start,
                   const int32_t *cell_boxes, int cell_cap, const uint32_t *cell4, const uint32_t *occ, const int32_t *chunks,
                   int n_chunks, void *hits, int hit_cap, int32_t *n_hits, int32_t *chunk_box_count, int max_boxes,
                   int32_t *overflow, void *stream);
/* frame_chunk_off (F+1) i64; writes box_total (n_boxes) i32 and offsets (n_boxes+1) i64 (exclusive). */
int al3d_crop_scan(const int64_t *box_off, const int64_t *frame_chunk_off, int n_frames, int64_t n_boxes,
                   int32_t *chunk_box_count, int max_boxes, int32_t *box_total, int64_t *offsets, void *stream);
/* out_idx (capacity) i32 point index within its frame, ascending per box; out_xyz (capacity,3) f32 copy
 * (may be NULL); out_xyz_global (capacity,3) f64 = poses[f] (4x4 row-major f64) applied to [x y z 1]
 * (may be NULL, needs poses). */
int al3d_crop_fill(const int64_t *box_off, int n_frames,
                   const int32_t *chunks, int n_chunks, const void *hits, int hit_cap, const int32_t *n_hits,
                   const int32_t *chunk_box_count, int max_boxes, const int64_t *offsets, const double *poses,
                   int64_t capacity, int32_t *out_idx, float *out_xyz, double *out_xyz_global, int32_t *overflow,
                   void *stream);
/* dense (N, n_boxes) u8 mask (pre-zeroed) of one frame from its index lists. */
int al3d_crop_dense_mask(const int32_t *idx, const int64_t *offsets, int n_boxes, uint8_t *mask, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Device-side track preparation (dataset-side merge / resample / canonical-frame transform).
 * Replaces the numpy part of STATICTRACK.__getitem__ (tools/static_model.py:541-547,569-570) and of
 * DYNAMICTRACK.__getitem__ (tools/dynamic_model.py:429-453,503-507).  float64 arithmetic like the reference,
 * rounded to float32 once on output.
 * ---------------------------------------------------------------------------------------------- */

/* out[b,j,0:3] = Rz(-heading_b) (inv_pose_b [p;1] - centre_b), p = src_xyz[choice[b,j]] (or the origin when
 * choice[b,j] < 0: the reference's zero-padded frames go through the same transform); with c_out == 4
 * out[b,j,3] = 0.1 * (j / time_block - time_center).  src_xyz (rows,3) f64 global frame; choice (bs,n_out)
 * i64 absolute row indices (the resample of :546-547 / :431-437, drawn by the caller); inv_pose (bs,16) f64
 * row-major; init_box (bs, box_stride) f64 with centre in columns 0..2 and heading in column heading_col;
 * out (bs, n_out, c_out) f32 point-major. */
int al3d_track_points_prep(const double *src_xyz, const int64_t *choice, int bs, int n_out, const double *inv_pose,
                           const double *init_box, int box_stride, int heading_col, int c_out, int time_block,
                           int time_center, float *out, void *stream);

/* box (bs, steps, 8) f64 global-frame [x y z l w h heading dt] -> out (bs, steps, 8) f32 step-major in the
 * vehicle frame of inv_pose, relative to the centre step (transform_box tools/dynamic_model.py:511-525, :452,
 * :506-507); init_box (bs, 8) f64 receives the centre step before the subtraction (:489), may be NULL. */
int al3d_boxseq_prep(const double *box, int bs, int steps, int center_step, const double *inv_pose, float *out,
                     double *init_box, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Track-level glue around the models (integer / index work exact; float64 arithmetic like the reference's numpy).
 * ---------------------------------------------------------------------------------------------- */

/* tools/trackData.py:25-45: regroup per-frame detections by tracking id.  ids (n_obs) i64 and frame_of_obs (n_obs) i32
 * list the observations in the reference's iteration order (frame by frame, boxes in frame order).  Tracks are numbered
 * in order of first appearance; track_obs (track_cap, n_frames) i32 holds each track's observation indices in frame
 * order (-1 padded), track_len (track_cap) i32, track_id (track_cap) i64, track_of_obs (n_obs) i32, n_tracks (1) i32.
 * error (1) i32: 1 = track_cap / frame index out of range, 2 = an id appears twice in one frame.
 * Scratch: hash_keys (hash_size) i64, hash_first / hash_rank (hash_size) i32 with hash_size a power of two >= 2*n_obs,
 * presence (track_cap * n_frames) i32. */
int al3d_track_regroup(const int64_t *ids, const int32_t *frame_of_obs, int n_obs, int n_frames, int track_cap,
                       int64_t *hash_keys, int32_t *hash_first, int32_t *hash_rank, int hash_size, int32_t *presence,
                       int32_t *track_of_obs, int64_t *track_id, int32_t *track_obs, int32_t *track_len,
                       int32_t *n_tracks, int32_t *error, void *stream);
/* tools/motionState.py:47-49: feat (n_tracks, 2) f64 = [ ||b_first - b_last||, ||var(b, axis=0)|| ] over the first
 * n_cols columns of boxes (n_obs, box_stride) f64.  (The reference slices a (L,1,7) array with [0, :3], which keeps
 * all 7 box columns: pass n_cols = 7 for parity, 3 for the centre-only feature the comment there intends.) */
int al3d_motion_features(const int32_t *track_obs, const int32_t *track_len, int n_tracks, int n_frames, const double *boxes,
                         int box_stride, int n_cols, double *feat, void *stream);
/* Training labels of STATICTRACK.__getitem__ (tools/static_model.py:549-566, tools/utils.py:53-67).  mask_label
 * (bs, n_out) f32 (may be NULL): the resampled points (same src_xyz / choice / inv_pose as al3d_track_points_prep, before
 * the canonical transform) tested in float64 against gt_planes (bs,6,4) f32 from al3d_crop_box_setup;  centre label =
 * gt_box[:, :3]; heading class / residual = angle2class(gt heading - init_heading, 12); size class / residual =
 * size2class(gt lwh).  float64 arithmetic, rounded to float32 once on output. */
int al3d_track_labels(const double *src_xyz, const int64_t *choice, int bs, int n_out, const double *inv_pose,
                      const float *gt_planes, const float *gt_box, const double *init_heading, float *mask_label,
                      float *center_label, int64_t *heading_cls, float *heading_res, int64_t *size_cls, float *size_res,
                      void *stream);
/* tools/static_eval.py:84-92: the refined box of track t (final_box (n_tracks,7) f32, in the vehicle frame of its best
 * frame) -> every observation of the track: out_boxes[obs] = transform_box(transform_box(box, best_pose[t]),
 * obs_inv_pose[obs]) in float64 (poses 4x4 row-major). */
int al3d_box_writeback(const float *final_box, const double *best_pose, const int32_t *track_obs, const int32_t *track_len,
                       int n_tracks, int n_frames, const double *obs_inv_pose, double *out_boxes, void *stream);

/* det <-> GT matching (det3d/datasets/waymo/waymo_common.py:173-188): for every detection (n_det,7) f32 [x y z l w h
 * heading] of frame det_frame[i], the GT box of that frame (gt (sum M_f,7), gt_off (F+1) i64) with the largest rotated
 * 3-D IoU (det3d/ops/iou3d_nms/iou3d_nms_utils.py:35-72): best_idx (index within the frame, -1 if the frame has no GT)
 * and best_iou.  The caller applies the 0.75 threshold.  float32; parity with the external pcdet kernel is unpinned
 * (SURVEY.md 8c): checked against an independent float64 restatement and analytic cases. */
int al3d_match_iou3d(const float *det, const int32_t *det_frame, int64_t n_det, const float *gt, const int64_t *gt_off,
                     int32_t *best_idx, float *best_iou, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Loss forward (evaluation / logging; no backward).  Replaces FrustumPointNetLossOneBoxEst.forward
 * tools/static_model.py:348-425 (each head set of ...TwoBoxEst :427-517; DynamicModelLoss
 * tools/dynamic_model.py:321-398).  out6 = [mask NLL, centre Huber(2), heading CE, size CE, heading-residual
 * Huber(1), size-residual Huber(1)], unweighted means; logits == NULL skips the mask term (out6[0] = 0).
 * mask_label (M) f32 0/1; class labels i64; partial_ws (n_partial) f32 scratch.
 * ---------------------------------------------------------------------------------------------- */
int al3d_loss_forward(const float *logits, const float *mask_label, int64_t M, const float *center,
                      const float *center_label, const float *heading_scores, const int64_t *heading_cls_label,
                      const float *heading_res_norm, const float *heading_res_label, const float *size_scores,
                      const int64_t *size_cls_label, const float *size_res_norm, const float *size_res_label, int bs,
                      float *partial_ws, int n_partial, float *out6, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Training step (BASELINE.json configs[4]; reference loop tools/static_train.py:65-90, tools/dynamic_train.py:37-133).
 * Activations are row-major (M, C) with M = bs*n rows (Conv1d(k=1) on (bs,C,n) == Linear on the rows).  The layer
 * GEMMs are al3d_linear_f32 (forward and, on W^T, the input gradient); everything else is below.  fp32 arithmetic,
 * fp64 finalisation of the long column sums, fixed reduction orders (bit-reproducible).  `ws`: caller scratch of
 * al3d_train_ws_floats(M, C) / al3d_wgrad_ws_floats(M, N, K) floats.
 * ---------------------------------------------------------------------------------------------- */
int al3d_train_ws_floats(int64_t M, int C);
int al3d_wgrad_ws_floats(int64_t M, int N, int K);
/* nn.BatchNorm1d in training mode + ReLU (+ Dropout multiplier): batch mean / biased variance over the M rows,
 * z = relu(gamma * (y - mean) * rstd + beta) * drop; running_mean / running_var (may be NULL) are updated with
 * `momentum` and the unbiased variance (tools/static_model.py:254-258,279-283).  drop (may be NULL) is addressed as
 * drop[(m / rows_per_group) * drop_sg + c * drop_sc + (m % rows_per_group) * drop_sr] (element strides), so a mask
 * drawn in the reference's (bs, C, n) layout is read in place (tools/static_model.py:264,293).  mean / rstd (C) are
 * outputs kept for the backward. */
int al3d_bn_train_forward(const float *y, int64_t M, int C, const float *gamma, const float *beta, float eps, float momentum,
                          float *running_mean, float *running_var, const float *drop, int64_t drop_sg, int64_t drop_sc,
                          int64_t drop_sr, int64_t rows_per_group, int relu, float *ws, float *mean, float *rstd, float *z,
                          void *stream);
/* Backward of the same: dz (gradient w.r.t. z) -> dy (may alias dz), dgamma, dbeta (C). */
int al3d_bn_train_backward(const float *dz, const float *y, int64_t M, int C, const float *gamma, const float *beta,
                           const float *mean, const float *rstd, const float *drop, int64_t drop_sg, int64_t drop_sc,
                           int64_t drop_sr, int64_t rows_per_group, int relu, float *ws, float *dgamma, float *dbeta,
                           float *dy, void *stream);
/* Column sums of x (M, C): per group of rows_per_group rows -> out (M / rows_per_group, C), or of the whole matrix
 * (rows_per_group <= 0 or >= M; needs ws) -> out (C).  Bias gradients and the per-object sums of the dconv1 backward. */
int al3d_group_colsum(const float *x, int64_t M, int C, int64_t rows_per_group, float *ws, float *out, void *stream);
/* torch.max over the n points of each object with its arg-max row (tools/static_model.py:284,334) and the backward
 * scatter dz[(g*n + arg[g,c]), c] = dg[g,c] into a zero-filled dz. */
int al3d_group_max_forward(const float *z, int64_t G, int64_t n, int C, float *g, int32_t *arg, void *stream);
int al3d_group_max_backward(const float *dg, const int32_t *arg, int64_t G, int64_t n, int C, float *dz_zeroed, void *stream);
/* Weight gradient dW (N, K; row stride lddw) (+)= dY^T . X with dY (M, N; ldy), X (M, K; ldx). */
int al3d_wgrad_f32(const float *dy, int64_t ldy, const float *x, int64_t ldx, int64_t M, int N, int K, float *ws,
                   float *dw, int64_t lddw, int accumulate, void *stream);
/* Split-precision tensor-core GEMMs on fp32 operands: the layer GEMMs of the training step (the Conv1d(k=1) / Linear
 * forward and input gradient of tools/static_model.py:241-339 under .train() / .backward()).
 *   C (M, N; ldc) (+)= A (M, K; lda) . B^T (+ bias[N] | rowbias[row / rows_per_group][N])
 * B is (N, K; ldb), or with b_transposed != 0 the (K, N; ldb) matrix whose transpose is meant (dgrad: the weight itself).
 * K % 64 == 0; N is 64, 128 or a multiple of 256; A and C rows 16-byte aligned.  ws: al3d_gemm_split_ws_bytes(N, K, parts)
 * bytes of device scratch (the packed weight image).  parts = 2 ("bf16x3"): every fp32 value is hi + lo in bf16 and a
 * product is a_hi*b_hi + a_lo*b_hi + a_hi*b_lo, ~1e-5 relative to an fp32 GEMM; parts = 3 ("bf16x6"): hi + mid + lo (all
 * 24 mantissa bits) and the six products of combined order <= 2, ~1e-7 relative (fp32-grade).  fp32 accumulation. */
int64_t al3d_gemm_split_ws_bytes(int N, int K, int parts);
int al3d_gemm_split_nt(const float *a, int64_t lda, int M, int K, const float *b, int64_t ldb, int b_transposed,
                       const float *bias, const float *rowbias, int rows_per_group, int N, int accumulate,
                       float *c, int64_t ldc, int parts, void *ws, void *stream);
/* The weight gradient on the tensor cores: C (N, K; ldc) (+)= A^T . B with A (M, N; lda) = dY and B (M, K; ldb) = the layer
 * input, reduced over the M rows (both operands staged MN-major, no transposition); per-CTA partial tiles are summed in a
 * fixed order (deterministic).  N % 8 == 0; K in 64..Kc (steps of 32) or a multiple of Kc, Kc = 256 (parts 2) / 128
 * (parts 3); ws: al3d_gemm_split_tn_ws_bytes(M, N, K, parts) bytes of device scratch. */
int64_t al3d_gemm_split_tn_ws_bytes(int64_t M, int N, int K, int parts);
int al3d_gemm_split_tn(const float *a, int64_t lda, const float *b, int64_t ldb, int64_t M, int N, int K, int parts,
                       void *ws, float *c, int64_t ldc, int accumulate, void *stream);
/* Gradient of sum_t w6[t] * term_t (terms of al3d_loss_forward, w6 six DEVICE floats) w.r.t. the logits -> dlogits
 * (M, 2) (skipped when logits == NULL) and w.r.t. the 39-wide head vector -> dbox (bs, 39) = [centre 3 | heading scores
 * 12 | normalised heading residuals 12 | size scores 3 | normalised size residuals 9] (tools/static_model.py:341-425). */
int al3d_loss_backward(const float *logits, const float *mask_label, int64_t M, const float *center, const float *center_label,
                       const float *heading_scores, const int64_t *heading_cls_label, const float *heading_res_norm,
                       const float *heading_res_label, const float *size_scores, const int64_t *size_cls_label,
                       const float *size_res_norm, const float *size_res_label, int bs, const float *w6, float *dlogits,
                       float *dbox, void *stream);
/* count_zeroed += number of points whose arg-max class equals the label (seg accuracy, tools/static_train.py:128-129). */
int al3d_seg_correct(const float *logits, const float *mask_label, int64_t M, unsigned long long *count_zeroed, void *stream);
/* One torch.optim.Adam step (amsgrad off) over a flat bucket of n parameters; grad is multiplied by grad_scale first
 * (1 / world size after the gradient all-reduce); weight_decay is the L2 form (tools/static_train.py:220). */
int al3d_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int step, float grad_scale, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Tensor-core (tcgen05 / TMEM, bf16 operands, fp32 accumulate) shared-MLP kernels.
 * Weights are BatchNorm-folded and packed by the caller (3dal_pytorch_b200/engine_bf16.py) into
 * 16 KB blocks of 128 rows x 64 K in the "KP" layout (K/8 planes of rows x 16 bytes), stored in the
 * order the kernel consumes them.
 * ---------------------------------------------------------------------------------------------- */

typedef struct al3d_chain_weights {
    int32_t c_in;            /* input channels (1..8)                                            */
    int32_t w0;              /* width of the first layer (CUDA cores): 64 or 128                  */
    int32_t n_mid;           /* number of chained tensor-core layers: 2 or 3                      */
    int32_t mid[3];          /* their widths (64 / 128 / 256)                                     */
    int32_t last;            /* width of the max-pooled last layer (512 / 1024)                   */
    int32_t n_blocks;        /* number of 16 KB blocks in wstream                                 */
    const float *w0_w;       /* (8, w0) fp32, transposed, zero rows for c >= c_in                */
    const float *w0_b;       /* (w0)                                                              */
    const float *mid_b;      /* concatenated fp32 biases of the mid layers                        */
    const float *last_b;     /* (last)                                                            */
    const void  *wstream;    /* packed bf16 weight blocks                                         */
} al3d_chain_weights;

/* first layer -> chained MMA layers -> last layer max-pooled over the n points of each object.
 * out (bs,last) must be zero-filled by the caller; results are relu(max + bias) >= 0.
 * Replaces ins_seg conv1-5 + torch.max (tools/static_model.py:279-284), the static box-head trunk
 * (:330-334) and the PointEmbedding / BoxEmbedding trunks (tools/dynamic_model.py:241-245,278-282). */
int al3d_chain_maxpool_bf16(const al3d_chain_weights *w, const float *x, int64_t sb, int64_t sc, int64_t sp,
                            int bs, int n, float *out, void *stream);

typedef struct al3d_pass1_weights {
    int32_t c_in;
    int32_t reserved;
    const float *w1_w, *w1_b;          /* ins_seg.conv1 folded fp32: (8,64) transposed + padded, (64)  */
    const float *b2, *b3, *b4, *b5;    /* conv2-5 biases (64),(64),(128),(1024)                         */
    const void  *wfront;               /* conv2, conv3, conv4 packed bf16, one 16 KB slot each          */
    const void  *w5stream;             /* conv5: 16 packed blocks of 128 channels x 64 K, (chunk, k-block) order */
    const float *consts_host;          /* HOST memory, 832 floats: w1_w (512) | w1_b (64) | b2 (64) | b3 (64) | b4 (128) --
                                          copied into the kernel parameter block (constant bank operands)          */
} al3d_pass1_weights;

/* First half of PointNetInstanceSeg.forward (tools/static_model.py:279-284): conv1..conv5 (+BN+ReLU) and the
 * max over the n points of each object -> out (bs,1024), which must be zero-filled by the caller.
 * Specialised, faster variant of al3d_chain_maxpool_bf16 for the segmentation widths (tile pairs, N = 256). */
int al3d_seg_pass1_bf16(const al3d_pass1_weights *w, const float *x, int64_t sb, int64_t sc, int64_t sp,
                        int bs, int n, float *out, void *stream);

typedef struct al3d_pass2_weights {
    int32_t c_in;
    int32_t reserved;
    const float *w1_w, *w1_b;      /* ins_seg.conv1 folded fp32: (8,64) transposed + padded, (64)  */
    const float *b2;               /* conv2 bias (64)                                               */
    const float *bd2, *bd3, *bd4;  /* dconv2-4 biases (256),(128),(128)                             */
    const float *w5, *b5;          /* dconv5 fp32 (2,128), (2)                                      */
    const void  *wstream;          /* two per-CTA halves of the 23 packed bf16 blocks (conv2, dconv1/dconv2 interleaved, dconv3, dconv4) */
} al3d_pass2_weights;

/* Second half of PointNetInstanceSeg.forward (tools/static_model.py:286-295) + the mask of
 * point_cloud_masking (:59).  gbias (bs,512) = W_dconv1[:, 64:] . global_feature + folded bias
 * (the concat-with-global-feature turned into a per-object bias).  Writes logits (bs,n,2) f32 and
 * mask (bs,n) u8. */
int al3d_seg_pass2_bf16(const al3d_pass2_weights *w, const float *x, int64_t sb, int64_t sc, int64_t sp,
                        int bs, int n, const float *gbias, float *logits, uint8_t *mask, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Split-precision ("bf16x3") tensor-core mode: the parity-grade variant of the three kernels above.  Every fp32
 * activation and weight is carried as hi = bf16(x), lo = bf16(x - hi) and every product is evaluated as
 * a_hi*w_hi + a_lo*w_hi + a_hi*w_lo (three tcgen05.mma per K step, fp32 accumulation): 16 significant bits per
 * operand, logits within ~5e-5 of the fp32 reference (bf16: ~2e-2, fp16 / tf32 operands: ~3e-3).
 * Weights: 16 KB slots, KP layout, a hi slot followed by a lo slot per (<=128 rows x 64 K) block, in consumption order.
 * "mixed" mode (last_f16 / d2_mode below): the two widest layers of PointNetInstanceSeg -- conv5 128 -> 1024
 * (tools/static_model.py:283) and dconv2 512 -> 256 (:290), 73 % of the network's MACs -- multiply IEEE fp16 operands
 * (11 significant bits, fp32 accumulation) with one / two MMAs per product instead of three; every other layer stays
 * bf16x3; the box-head / embedding trunks may run their max-pooled last layer the same way (last_f16 with pair = 0).
 * Logits within ~4e-4 of the fp32 reference (profiles/r2_precision_study_mixed.txt).  Activations above the fp16 range
 * (65504) saturate in those layers.
 * ---------------------------------------------------------------------------------------------- */
typedef struct al3d_split_chain_weights {
    int32_t c_in;            /* input channels (1..8)                                                          */
    int32_t w0;              /* width of the first layer (CUDA cores): 64 or 128                                */
    int32_t n_mid;           /* number of chained tensor-core layers: 2 or 3                                    */
    int32_t mid[3];          /* their widths (64 / 128 / 256)                                                   */
    int32_t last;            /* width of the max-pooled last layer (multiple of 128, <= 1024)                   */
    int32_t n_blocks;        /* number of 16 KB slots in wstream                                                */
    int32_t pair;            /* 1: the last layer runs on pairs of 128-point tiles (its input must fit 128 KB)  */
    int32_t last_f16;        /* 1: the max-pooled last layer multiplies IEEE fp16 operands, ONE MMA per product (its
                                blocks are single fp16 slots, no lo slot): the "mixed" mode, see above               */
    const float *w0_w;       /* (8, w0) fp32, transposed, zero rows for c >= c_in                               */
    const float *w0_b;       /* (w0)                                                                            */
    const float *mid_b;      /* concatenated fp32 biases of the mid layers                                      */
    const float *last_b;     /* (last)                                                                          */
    const void  *wstream;    /* mid layers in (layer, row chunk, k block) order, then the last layer (chunk, k block) */
} al3d_split_chain_weights;
/* Same contract as al3d_chain_maxpool_bf16 (ins_seg conv1-5 + max with pair = 1; the three head / embedding trunks). */
int al3d_chain_maxpool_bf16x3(const al3d_split_chain_weights *w, const float *x, int64_t sb, int64_t sc, int64_t sp,
                              int bs, int n, float *out, void *stream);

typedef struct al3d_split_tail_weights {
    int32_t c_in;
    int32_t d2_mode;               /* dconv2 (512 -> 256, 61 % of this kernel's MACs): 0 = bf16x3 like every other layer (54 slots);
                                      1 = fp16 activations x fp16 weights, one MMA per product; 2 = fp16 hi + lo activations x fp16
                                      weights, two MMAs.  1 / 2: the sixteen p(c) blocks are single fp16 slots (38 slots in all)   */
    const float *w1_w, *w1_b;      /* ins_seg.conv1 folded fp32: (8,64) transposed + padded, (64)  */
    const float *b2;               /* conv2 bias (64)                                               */
    const float *bd2, *bd3, *bd4;  /* dconv2-4 biases (256),(128),(128)                             */
    const float *w5, *b5;          /* dconv5 fp32 (2,128), (2)                                      */
    const void  *wstream;          /* 54 slots: conv2 | d1(0) d1(1) d1(2) p(0) d1(3) p(1) p(2) p(3) | dconv3 | dconv4 (csrc/chain_split.cu) */
    const void  *wstream_pair;     /* NULL, or the same 54 slots as two per-CTA images of 8 KB half slots (CTA r: rows
                                      [r*R/2, (r+1)*R/2) of every block): selects the CTA-pair (cta_group::2) kernel */
} al3d_split_tail_weights;
/* Same contract as al3d_seg_pass2_bf16. */
int al3d_seg_pass2_bf16x3(const al3d_split_tail_weights *w, const float *x, int64_t sb, int64_t sc, int64_t sp,
                          int bs, int n, const float *gbias, float *logits, uint8_t *mask, void *stream);

/* D(128,N) fp32 = A(128,K) . B(N,K)^T from KP-packed bf16 operands with one tcgen05.mma chain:
 * unit test of the descriptor / layout conventions. */
int al3d_umma_selftest(const void *a_kp, const void *b_kp, int N, int K, float *d_out, int swap_lbo_sbo, void *stream);
/* Same product with A (128,K) fp32 row-major packed to bf16 and staged in TMEM by the kernel (A-from-TMEM MMA). */
int al3d_umma_selftest_ts(const float *a, const void *b_kp, int N, int K, float *d_out, void *stream);
/* CTA-pair (cluster of 2, cta_group::2) variant: a (256,K) fp32, b_kp_halves = two KP-packed halves of B (N/2 rows each),
 * d_out (256,N).  Checks the 2-SM MMA conventions. */
int al3d_umma_selftest_pair(const float *a, const void *b_kp_halves, int N, int K, float *d_out, void *stream);
/* CTA-pair variant with both operands in shared memory, N = 128: a_kp_halves = two KP tiles of 128 rows, b_kp_halves =
 * two KP tiles of 64 rows (staged by the kernel as a sub-tile of a 128-row tile).  d_out (256,128). */
int al3d_umma_selftest_pair_ss(const void *a_kp_halves, const void *b_kp_halves, int K, float *d_out, void *stream);

/* Development aid: issue n_mma tcgen05.mma (M = 128, N, K = 16; mode 0 = both operands in shared memory, 1 = A from
 * TMEM) from one thread on each of n_ctas CTAs, a commit every commit_every MMAs (0 = only at the end).
 * background: bit 0 = four warps read the accumulator columns with tcgen05.ld meanwhile, bit 1 = one thread streams
 * 16 KB blocks from src_1mib (device, >= 1 MiB) into shared memory meanwhile, bit 2 = the issuing lane is chosen with
 * elect.sync instead of `lane == 0`.
 * out (6 int64, CTA 0): issue cycles, cycles to completion, issue cycles of the first 8, background blocks, background
 * loads, scratch.  scripts/mma_microbench.py. */
int al3d_mma_microbench(int N, int n_mma, int commit_every, int mode, int n_ctas, int background, const void *src_1mib,
                        long long *out, void *stream);

/* Watchdog of the tensor-core kernels.  Every mbarrier wait in them is bounded; a wait that gives up (a protocol
 * bug) records a code in the status word of the device the kernel runs on and traps, so the launch fails with a
 * CUDA error at the caller's next synchronisation instead of returning invalid outputs.  The status word lives in
 * pinned host memory mapped into the device (one block per device, allocated at the first tensor-core launch):
 * al3d_tc_abort_code reads and clears the CURRENT device's word with a plain host load -- no CUDA call, no
 * synchronisation; it is meaningful for launches the caller has already synchronised with.
 * al3d_tc_status_word_host returns the word's host address for callers that poll it themselves. */
int al3d_tc_abort_code(int *code_host);
int al3d_tc_status_word_host(const void **word_host);
/* Diagnostics: trap_on_timeout = 0 records the code without trapping; stress_ns > 0 makes every role of the
 * tensor-core kernels sleep a pseudo-random time (< stress_ns) before its barrier waits, to shake out protocol
 * races (tests/test_gpu_stress.py).  Defaults: 1, 0 (environment AL3D_TC_TRAP / AL3D_TC_STRESS_NS). */
int al3d_tc_configure(int trap_on_timeout, int stress_ns);

/* Development aid: when set to a device buffer of 2*3*4*64 int64, CTA 0 of al3d_seg_pass2_bf16 (first half)
 * and of al3d_seg_pass1_bf16 (second half) record a clock64 timeline of their first four tiles / tile pairs
 * (role-major: MMA thread, epilogue thread, producer).  Pass
 * NULL to switch it off (the default). */
int al3d_set_debug_buffer(void *dev_ptr);

#ifdef __cplusplus
}
#endif
#endif /* AL3D_H_ */
