/*
 * moseq_b200.h -- C ABI of libmoseq_b200.so: the B200 (sm_100a) implementation of the per-frame
 * extract hot path of tischfieldlab/moseq2-detectron-extract.
 *
 * The reference has NO FFI for this path: it is pure Python that calls NumPy/OpenCV
 * (SURVEY.md section 8b).  Each entry point below therefore replaces one reference *Python function*
 * (cited as `ref: file:line` relative to /root/reference/moseq2_detectron_extract/) and is what a
 * ctypes stub in that function's place would bind (INTEGRATION.md shows the stubs).
 *
 * Conventions
 *   - plain pointers + sizes, no torch / C++ types; every image is dense row-major
 *   - "dev" pointers are CUDA device pointers on the current device; "host" pointers are host memory
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream)
 *   - device entry points never allocate device memory, never synchronise with the host and are re-entrant per
 *     stream (msq_extract_chunk keeps one msq_engine -- a side stream + two events -- per host thread and device; callers
 *     that want no library-held state create their own with msq_engine_create and call msq_extract_chunk_engine);
 *     scratch memory is passed in by the caller (sizes from the *_scratch_bytes helpers)
 *   - return 0 on success, a negative MSQ_E* code otherwise; msq_last_error() gives the message of
 *     the last failure on the calling thread (Python raises from it)
 *   - frames are counted in "n"; per-frame outputs are SoA arrays of length n
 */
#ifndef MOSEQ_B200_H
#define MOSEQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MSQ_API __attribute__((visibility("default")))
#else
#define MSQ_API
#endif

#define MSQ_VERSION 100            /* major*10000 + minor*100 + patch */

#define MSQ_OK            0
#define MSQ_EINVAL      (-1)       /* bad argument (null pointer, non-positive size, bad enum) */
#define MSQ_ECUDA       (-2)       /* a CUDA runtime call or kernel launch failed */
#define MSQ_EUNSUPPORTED (-3)      /* valid request this build cannot serve (e.g. ROI wider than 4096 px) */
#define MSQ_ENOMEM      (-4)       /* scratch buffer too small / allocation failed (engine only) */

/* dtype codes for the background image (ref: SURVEY trap 8: float64 fresh, uint16 from TIFF cache) */
#define MSQ_BG_NONE 0
#define MSQ_BG_F32  1
#define MSQ_BG_F64  2
#define MSQ_BG_U16  3

/* flags of msq_prep_frames */
#define MSQ_PREP_HAS_VMIN 1
#define MSQ_PREP_HAS_VMAX 2

#define MSQ_NUM_KEYPOINTS 8        /* ref: io/annot.py:51-60 */
#define MSQ_NUM_SCALARS  17        /* ref: proc/scalars.py:13-31 */
#define MSQ_NUM_KPT_COLS 96        /* ref: proc/keypoints.py:147-163: 8 kpts x 2 systems x 6 */

MSQ_API int         msq_version(void);
MSQ_API const char *msq_last_error(void);
/* number of SMs / name of the current device (host-side launch sizing, diagnostics) */
MSQ_API int         msq_device_info(int *sm_count, int *cc_major, int *cc_minor, char *name, int name_len);

/* per-kernel device timing with CUDA events recorded on the launching stream (off by default; the
 * equivalent of the reference's MOSEQ_DETECTRON_PROFILE switch, ref: io/util.py:239-255).  collect()
 * synchronises on the recorded events and ADDS into total_ms[k] / timed[k] (k < msq_kernel_count()). */
MSQ_API int msq_kernel_timing_enable(int enable);
MSQ_API int msq_kernel_timing_collect(double *total_ms, long long *timed, int capacity);
MSQ_API int msq_kernel_count(void);
MSQ_API const char *msq_kernel_name(int id);
MSQ_API long long msq_kernel_launches(int id);

/* ---- a2  prep_raw_frames  (ref: proc/proc.py:129-172; apply_roi proc/roi.py:215-236; get_bbox :239-254;
 *                            find_invalid_pixels proc/proc.py:175-186) ---------------------------------
 * frames_dev  (n,H,W) int16 raw depth      bground_dev (H,W) of bg_dtype or NULL/MSQ_BG_NONE
 * roi_dev     (H,W) uint8, nonzero = inside, or NULL (then the box must be the full frame)
 * box         y0,x0 + h,w : the ROI bounding box, max-exclusive as the reference slices it
 * out_dev     (n,h,w) uint8: ((bground - frame) * roi)[box], <vmin -> 0, >vmax -> vmax, truncated
 * invalid_count_dev (n) int32 or NULL: number of raw==0 pixels inside roi&box per frame (the pixels the
 *             reference in-paints, proc/proc.py:189-210)
 * invalid_bits_dev (n, h, ceil(w/8)) bytes or NULL, 4-byte aligned and padded to a multiple of 4 bytes: the same
 *             pixels as a packed mask (bit b of byte B of a row = pixel 8B+b), the input of msq_inpaint_frames */
MSQ_API int msq_prep_frames(const int16_t *frames_dev, int n, int H, int W,
                    const void *bground_dev, int bg_dtype, const uint8_t *roi_dev,
                    int y0, int x0, int h, int w, double vmin, double vmax, int flags,
                    uint8_t *out_dev, int32_t *invalid_count_dev, uint8_t *invalid_bits_dev, void *stream);
/* The same with one more output for the cleaning pass (prep -> clean fusion of the part that can be fused: the prepared frame
 * itself has to exist, the R-CNN, the masked sums and the crops read it): positive_bits_dev (n, h, ceil(w/32)) uint32 or NULL,
 * msq_positive_bits_bytes(n,h,w) bytes, bit b of word i of a row = prepared pixel 32i+b is > 0.  msq_extract_chunk_engine /
 * msq_clean_frames_ws find the rows the 9x9 opening can leave non-zero from these 1-bit rows instead of re-reading and
 * re-thresholding the 8-bit frame.  Any SUPERSET of the positive pixels keeps the result exact (so the rows stay usable after
 * msq_inpaint_frames, which only rewrites pixels that were positive); a caller that edits the frames otherwise passes NULL. */
MSQ_API size_t msq_positive_bits_bytes(int n, int h, int w);
MSQ_API int msq_prep_frames_bits(const int16_t *frames_dev, int n, int H, int W,
                    const void *bground_dev, int bg_dtype, const uint8_t *roi_dev,
                    int y0, int x0, int h, int w, double vmin, double vmax, int flags,
                    uint8_t *out_dev, int32_t *invalid_count_dev, uint8_t *invalid_bits_dev, uint32_t *positive_bits_dev,
                    void *stream);

/* Instance masks handed over from the host as bit rows (the reference passes bool images, ref: proc/proc.py:672-684; one bit per
 * pixel is the same information in 1/8 of the PCIe bytes): bits_dev (n, h, ceil(w/8)), bit b of byte B of a row = pixel 8B+b
 * (numpy.packbits(..., axis=-1, bitorder='little')) -> out_dev (n,h,w) u8 in {0,1}, the mask format of every entry point. */
MSQ_API int msq_unpack_mask_bits(const uint8_t *bits_dev, int n, int h, int w, uint8_t *out_dev, void *stream);

/* Host -> device staging of the ROI box only, as ONE strided DMA transfer per chunk (host side: no kernel): frames_host
 * (n,H,W) int16 in pinned host memory -> out_dev (n,h,w) int16 = rows [y0,y0+h) x columns [x0,x0+w) of every frame.  Feed the
 * result to msq_prep_frames with H = h, W = w, y0 = x0 = 0 and the background / ROI cropped to the same box. */
MSQ_API int msq_copy_roi_rows(const int16_t *frames_host, int n, int H, int W, int y0, int x0, int h, int w, int16_t *out_dev,
                      void *stream);
/* The same for a ROI that fills only part of its box (the bucket floor is a disc): n_bands horizontal bands, band b = rows
 * [band_y[b], band_y[b+1]) x columns [band_x0[b], band_x1[b]) relative to the box (HOST int arrays; band_y has n_bands + 1
 * entries), one strided DMA transfer each into the same dense (n,h,w) array.  Every band must cover the ROI pixels of its rows;
 * what lies outside the bands is not written and never reaches msq_prep_frames' output (it multiplies by the ROI mask) -- keep
 * out_dev zero-initialised.  16 bands of a disc: 15 % fewer PCIe bytes than the box. */
MSQ_API int msq_copy_roi_bands(const int16_t *frames_host, int n, int H, int W, int y0, int x0, int h, int w, const int *band_y,
                       const int *band_x0, const int *band_x1, int n_bands, int16_t *out_dev, void *stream);

/* ---- a2  fill_invalid_pixels (ref: proc/proc.py:189-210): cv2.inpaint(frame, mask, radius, INPAINT_NS), bit-exact.
 * frames_dev (n_total,h,w) u8 updated IN PLACE; invalid_bits_dev as written by msq_prep_frames; frame_idx_dev (m) int32
 * indices of the frames to in-paint (NULL = frames 0..m-1); radius 1..4 (the reference uses 3).
 * scratch_dev: msq_inpaint_scratch_bytes(m,h,w) bytes, 8-byte aligned (per-frame distance map + spill space of the
 * priority queue). One warp per frame runs OpenCV's fast-marching order exactly. */
MSQ_API size_t msq_inpaint_scratch_bytes(int m, int h, int w);
MSQ_API int msq_inpaint_frames(uint8_t *frames_dev, const uint8_t *invalid_bits_dev, const int32_t *frame_idx_dev,
                       int m, int h, int w, int radius, void *scratch_dev, size_t scratch_bytes, void *stream);

/* ---- a3  scale_raw_frames (ref: proc/proc.py:214-234, dtype uint8) --------------------------------------
 * out = trunc((in - vmin) * (255 / (vmax - vmin)) + 0); vmin_is_int selects NumPy's uint8 wrap-around
 * subtraction that applies when the reference is handed a Python int vmin. in/out may alias. */
MSQ_API int msq_scale_frames(const uint8_t *in_dev, uint8_t *out_dev, size_t count,
                     double vmin, double vmax, int vmin_is_int, void *stream);
/* same map, written as 3 identical channel planes (n,3,h,w) of float32 for the R-CNN input
 * (ref: model/predict.py:74-90 replicates the grey channel and moves it to CHW) */
MSQ_API int msq_scale_frames_chw3_f32(const uint8_t *in_dev, float *out_dev, int n, int h, int w,
                              double vmin, double vmax, int vmin_is_int, void *stream);

/* a3 fused with the detector's image transform (ref: model/predict.py:77-98 staging + the GeneralizedRCNNTransform of the
 * graph): out_dev (n, ph, pw, 3) CHANNELS-LAST bf16 (out_is_bf16 != 0) or float32 = zero-padded canvas whose top-left
 * (oh, ow) block is bilinear_resize((scale(in) - mean[c]) / std[c]) -- scale() as msq_scale_frames, resize as
 * torch.nn.functional.interpolate(mode="bilinear", align_corners=False) from (h, w) to (oh, ow); mean / std: 3 floats in
 * HOST memory, in the 0..255 units of the scaled image. */
MSQ_API int msq_detector_input(const uint8_t *in_dev, void *out_dev, int out_is_bf16, int n, int h, int w, int oh, int ow,
                       int ph, int pw, const float *mean_host, const float *std_host, double vmin, double vmax,
                       int vmin_is_int, void *stream);

/* The detector's stem in one kernel (ref: model/predict.py:74-77 grey -> 3 identical channels; detectron2 BasicStem of the
 * graph ref model/config.py configures): a3 scaling -> (x - mean) / std -> zero padding to (ph, pw) -> 7x7 stride-2 convolution
 * + bias -> ReLU -> 3x3 stride-2 max-pool.  Valid when PIXEL_MEAN / PIXEL_STD are the same for the three channels: the
 * 3-channel convolution of a replicated grey image is the 1-channel convolution with the weights summed over the input
 * channels.  in_dev (n,h,w) u8 prepared frames; w49x64_dev (49, 64) float32 = sum over input channels of the (folded) stem
 * weight, [tap r*7+s][output channel]; bias64_dev (64) float32; out_dev (n, PH, PW, 64) channels-last bf16 / float32 with
 * PH = ((ph-1)/2+1 - 1)/2 + 1 (64 for ph = 256). */
MSQ_API int msq_stem_conv_pool(const uint8_t *in_dev, int n, int h, int w, int ph, int pw, double vmin, double vmax, int vmin_is_int,
                       float mean, float std, const float *w49x64_dev, const float *bias64_dev, void *out_dev, int out_is_bf16,
                       void *stream);

/* The same stem on the tensor cores (csrc/stem_tc.cu; bf16 output): the 49 taps are one 64-wide K block, the im2col operand is
 * written into shared memory in the swizzle-128B layout and multiplied by tcgen05.mma.  b_tile_dev: 8192 bytes = the summed stem
 * weight as bf16 [64 output channels][64 k] in that layout: element (n, k) at byte n*128 + (((k>>3) ^ (n&7))<<4) + (k&7)*2,
 * k = r*7+s < 49, zero beyond (model/rcnn.py stem_b_tile builds it). */
MSQ_API int msq_stem_conv_pool_tc(const uint8_t *in_dev, int n, int h, int w, int ph, int pw, double vmin, double vmax, int vmin_is_int,
                          float mean, float std, const void *b_tile_dev, const float *bias64_dev, void *out_dev, void *stream);

/* Segmented greedy NMS for the RPN proposal filtering of a whole batch (replaces the per-image box_ops.batched_nms calls of
 * torchvision's RegionProposalNetwork.filter_proposals behind Predictor; the reference's detectron2 RPN does the same
 * per-image loop).  boxes_dev (n,K,4) float32, per image sorted by descending score and already shifted per pyramid level
 * (the "coordinate trick"); valid_dev (n,K) u8; keep_dev (n,max_keep) int32 receives the indices of the first max_keep
 * survivors (-1 padded), count_dev (n) their number.  boxes_dev 16-byte aligned. */
MSQ_API int msq_nms_sorted(const float *boxes_dev, const uint8_t *valid_dev, int n, int K, float iou_threshold, int max_keep,
                   int32_t *keep_dev, int32_t *count_dev, void *stream);

/* The same greedy NMS for long keep lists (detectron2's RPN keeps the first POST_NMS_TOPK_TEST = 1000 survivors of ~3000
 * candidates per image): the K x K overlap matrix is evaluated in parallel as bit rows, then one warp per image walks the
 * candidates in order.  Same arguments and results as msq_nms_sorted; K <= 6144; scratch_dev: msq_nms_scratch_bytes(n, K)
 * bytes, 8-byte aligned. */
MSQ_API size_t msq_nms_scratch_bytes(int n, int K);
MSQ_API int msq_nms_sorted_long(const float *boxes_dev, const uint8_t *valid_dev, int n, int K, float iou_threshold, int max_keep,
                        int32_t *keep_dev, int32_t *count_dev, void *scratch_dev, size_t scratch_bytes, void *stream);
/* The same when the candidates carry their pyramid level: level_valid_dev[i] = 0 (not a candidate) or 1 + level, as
 * msq_rpn_select writes it.  detectron2's batched_nms never lets levels suppress each other, so the overlap bits are computed per
 * level (3.4x fewer pairs for 3 x 1000 + 192 + 48 candidates) and every (image, level) is walked by its own warp; at most
 * max_per_level (<= 2048) candidates per level.  scratch_dev: msq_nms_levels_scratch_bytes(...) bytes, 256-byte aligned. */
MSQ_API size_t msq_nms_levels_scratch_bytes(int n, int K, int n_levels, int max_per_level);
MSQ_API int msq_nms_levels_long(const float *boxes_dev, const uint8_t *level_valid_dev, int n, int K, int n_levels, int max_per_level,
                        float iou_threshold, int max_keep, int32_t *keep_dev, int32_t *count_dev, void *scratch_dev,
                        size_t scratch_bytes, void *stream);

/* detectron2 find_top_rpn_proposals up to the NMS for a batch of equally sized images, one launch: per-level top-k of the
 * objectness logits (exact k-th largest by radix select; ties in anchor order), Box2BoxTransform decoding of the selected
 * anchors (weights 1, anchors from the level's stride and its three cell anchors), clipping to (img_h, img_w), validity
 * (finite, non-empty), one sort of all levels' candidates by descending logit (invalid last, ties in (level, anchor) order).
 * pred_dev[l]: HOST array of device pointers to (n, H_l, W_l, 16) channels-last head outputs of bf16 / float32 -- channels
 * [3 logits, 3 x 4 deltas, 1 padding]; heights / widths / strides: HOST int arrays; cell_anchors: HOST float array
 * (n_levels, 3, 4).  Outputs, K = sum over levels of min(pre_topk, 3 H_l W_l) <= 4096 candidates per image in sorted order:
 * boxes_dev (n,K,4), shifted_dev (n,K,4) = boxes + level * (largest valid coordinate of the image + 1) (torchvision's
 * batched_nms coordinate trick), scores_dev (n,K) (-inf where invalid), valid_dev (n,K) u8: the inputs of msq_nms_sorted*. */
MSQ_API int msq_rpn_select(const void *const *pred_dev, const int *heights, const int *widths, const int *strides,
                   const float *cell_anchors, int n_levels, int is_bf16, int n, int pre_topk, int img_h, int img_w,
                   float *boxes_dev, float *shifted_dev, float *scores_dev, uint8_t *valid_dev, void *stream);

/* Keypoint decoding for a batch (replaces the per-RoI loop of torchvision's heatmaps_to_keypoints / detectron2's keypoint
 * head inference): for every RoI and keypoint, the arg-max of the heatmap (maps_dev (R,K,Hm,Wm) float32) bicubically
 * resized to the RoI's ceil(width) x ceil(height), mapped back to image coordinates.  rois_dev (R,4) float32 x1,y1,x2,y2;
 * xyv_dev (R,K,3) float32 (x, y, 1); scores_dev (R,K) the heatmap value at the arg-max.  round_bf16 != 0: values are rounded
 * to bfloat16 before the comparison (what the resize yields under bf16 autocast). */
MSQ_API int msq_keypoints_from_heatmaps(const float *maps_dev, const float *rois_dev, int n_rois, int K, int Hm, int Wm,
                                int round_bf16, float *xyv_dev, float *scores_dev, void *stream);

/* Multi-level RoIAlign (torchvision.ops.MultiScaleRoIAlign / ops/poolers.py:_multiscale_roi_align, aligned = false) in one
 * launch.  feat_dev[l]: HOST array of n_levels device pointers to CHANNELS-LAST feature maps (n, H_l, W_l, C) of bf16
 * (is_bf16 != 0) or float32; heights / widths / scales: HOST arrays per level.  rois_dev (n_rois,5) float32 = (image
 * index, x1, y1, x2, y2) in image coordinates; levels_dev (n_rois) int64 = pyramid level of every RoI (LevelMapper; may
 * be NULL when n_levels == 1).  out_dev (n_rois, C, P, P) in the dtype of the features.  C % 8 == 0, sampling_ratio 1..4. */
MSQ_API int msq_roi_align_levels(const void *const *feat_dev, const int *heights, const int *widths, const float *scales,
                         int n_levels, int C, int is_bf16, const float *rois_dev, const long long *levels_dev, int n_rois,
                         int P, int sampling_ratio, void *out_dev, void *stream);

/* ---- a4  glue kernels of the repo's own detectron2-configuration graph (model/rcnn.py; ref: model/config.py:21-94 on top of
 *      detectron2's COCO-Keypoints/keypoint_rcnn_R_50_FPN_3x.yaml; export contract ref: model/deploy.py:65-110) -----------
 * GroupNorm of the FPN convolutions (FPN.NORM = 'GN', ref: model/config.py:82) on a channels-last map (n,H,W,C) of bf16
 * (is_bf16 != 0) or float32, fused with the top-down path (FPN.FUSE_TYPE = 'avg', ref: model/config.py:83):
 *   out = (GroupNorm(x; groups, eps, gamma, beta) + nearest_upsample_x2(top)) * scale        top (n,ceil(H/2),ceil(W/2),C) or NULL
 * gamma / beta: (C) float32 on the device.  Channels per group must be a multiple of 8.  x and out may alias.
 * scratch_dev: msq_group_norm_scratch_bytes(n,H,W,C) bytes, 16-byte aligned. */
MSQ_API size_t msq_group_norm_scratch_bytes(int n, int H, int W, int C);
MSQ_API int msq_group_norm_nhwc(const void *x_dev, int is_bf16, int n, int H, int W, int C, int groups, float eps,
                        const float *gamma_dev, const float *beta_dev, const void *top_dev, float scale, void *out_dev,
                        void *scratch_dev, size_t scratch_bytes, void *stream);

/* detectron2 ROIPooler with ROIAlignV2 (torchvision.ops.roi_align(aligned=True); sampling_ratio 0 = adaptive grid
 * ceil(roi / P), the detectron2 default) over n_levels channels-last maps (n, H_l, W_l, C), level of a box =
 * clamp(floor(canonical_level + log2(sqrt(area) / canonical_size + 1e-8)), min_level, min_level + n_levels - 1) as in
 * detectron2.modeling.poolers.assign_boxes_to_levels.  boxes_dev (n_rois,4) float32 x1,y1,x2,y2; box r belongs to image
 * r / rois_per_image.  out_dev (n_rois, P, P, C) CHANNELS-LAST in the dtype of the maps.  Host arrays as in
 * msq_roi_align_levels. */
MSQ_API int msq_roi_align_v2(const void *const *feat_dev, const int *heights, const int *widths, const float *scales, int n_levels,
                     int C, int is_bf16, const float *boxes_dev, int n_rois, int rois_per_image, int P, int sampling_ratio,
                     int min_level, int canonical_level, float canonical_size, void *out_dev, void *stream);

/* detectron2 fast_rcnn_inference with TEST.DETECTIONS_PER_IMAGE = 1 (ref: model/config.py:75) and one foreground class: per
 * image the proposal with the best soft-max foreground score above score_thresh (it always survives its class's NMS),
 * decoded with Box2BoxTransform(weights4_host) and clipped to the image.  pred_dev (n*k, pred_stride) float32 rows
 * [logit_fg, logit_bg, dx, dy, dw, dh, ...]; proposals_dev (n,k,4); counts_dev (n) int32 valid proposals per image or NULL.
 * box_dev (n,4), score_dev (n), has_dev (n) u8 (0: no detection, box = 0), index_dev (n) int32 or NULL. */
MSQ_API int msq_fastrcnn_top1(const float *pred_dev, int pred_stride, const float *proposals_dev, const int32_t *counts_dev, int n,
                      int k, int img_h, int img_w, float score_thresh, const float *weights4_host, float *box_dev,
                      float *score_dev, uint8_t *has_dev, int32_t *index_dev, void *stream);

/* The last layer of detectron2's KRCNNConvDeconvUpsampleHead (the reference's keypoint head, model/config.py:84 -> ROI_KEYPOINT_HEAD
 * defaults): F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False) of the deconvolution output.  in_dev
 * (n,K,H,W) bf16 (in_is_bf16) or fp32 with element strides sN,sC,sH,sW (channels-last or dense) -> out_dev (n,K,2H,2W) fp32 dense. */
MSQ_API int msq_upsample2x_bilinear(const void *in_dev, int in_is_bf16, long long sN, long long sC, long long sH, long long sW, int n, int K,
                            int H, int W, float *out_dev, void *stream);
/* detectron2.structures.keypoints.heatmaps_to_keypoints for a batch: as msq_keypoints_from_heatmaps, but the third column
 * of xyp_dev (R,K,3) is detectron2's score exp(max) / sum(exp(pool-resolution map)); logit_dev (R,K) or NULL receives the
 * heatmap value at the arg-max. */
MSQ_API int msq_keypoints_from_heatmaps_d2(const float *maps_dev, const float *rois_dev, int n_rois, int K, int Hm, int Wm,
                                   float *xyp_dev, float *logit_dev, void *stream);

/* The graph's 1x1 / 3x3 convolutions and Linear layers on the 5th-generation tensor cores (csrc/conv_tc.cu): implicit GEMM,
 * TMA-fed (the 3x3 taps are shifted reads of one 4-D tensor map, out-of-image rows / columns zero-filled by the TMA unit),
 * tcgen05.mma with fp32 accumulators in TMEM, fused epilogue out = act(conv + bias + residual).
 * x_dev (n,H,W,cin) channels-last bf16; w_dev (cout,k,k,cin) bf16; bias_dev (cout) float32 / bf16 (bias_is_bf16) or NULL; residual_dev like out or
 * NULL; out_dev (n,Ho,Wo,cout) bf16, Ho = (H-1)/stride + 1.  k = 1 (stride 1 or 2) or k = 3 (stride 1, padding 1); cin, cout
 * multiples of 64.  A Linear layer y = x W^T + b is the call with n = H = 1, W = rows. */
MSQ_API int msq_conv_tc(const void *x_dev, int n, int H, int W, int cin, const void *w_dev, int cout, int ksize, int stride,
                const void *bias_dev, int bias_is_bf16, const void *residual_dev, int relu, void *out_dev, void *stream);

/* ---- a6  clean_frames(iters_tail=3) (ref: proc/proc.py:480-515) -------------------------------------
 * 3x3 median (replicate border) then ONE opening with the 9x9 ellipse (SURVEY trap 3). in != out. */
MSQ_API int msq_clean_frames(const uint8_t *in_dev, uint8_t *out_dev, int n, int h, int w, void *stream);
/* The same with scratch memory (msq_clean_scratch_bytes(n,h,w) bytes, 8-byte aligned): the row pre-pass that finds where the
 * opening can be non-zero runs as its own lean launch instead of inside the pipeline kernel and also writes the zero rows, and
 * the pipeline kernel gets the remaining rows in equal shares (faster; msq_extract_chunk uses it).  positive_bits_dev: the bit
 * rows msq_prep_frames_bits wrote for in_dev, or NULL (then the pre-pass thresholds in_dev itself). */
MSQ_API size_t msq_clean_scratch_bytes(int n, int h, int w);
MSQ_API int msq_clean_frames_ws(const uint8_t *in_dev, const uint32_t *positive_bits_dev, uint8_t *out_dev, int n, int h, int w,
                        void *scratch_dev, size_t scratch_bytes, void *stream);

/* ---- a7  get_frame_features + im_moment_features (ref: proc/proc.py:237-302, 518-549) ----------------
 * fm = (cleaned > frame_threshold) & (mask != 0); polygon moments of the largest outer contour.
 * centroid_dev (n,2) f64 [x,y]; orientation_dev (n) f64 radians; axis_length_dev (n,2) f64;
 * sums24_dev (n,6) int64 or NULL: exact 24x polygon integrals (1,x,y,xx,xy,yy) of the winning blob.
 * scratch_dev: msq_frame_features_scratch_bytes(n,h,w) bytes, 4-byte aligned (the list of frames the streaming
 * row-convex fast path passes on to the general flood/peel kernel); NULL or too small: every frame takes the general kernel. */
MSQ_API size_t msq_frame_features_scratch_bytes(int n, int h, int w);
MSQ_API int msq_frame_features(const uint8_t *cleaned_dev, const uint8_t *mask_dev, int n, int h, int w,
                       double frame_threshold, double *centroid_dev, double *orientation_dev,
                       double *axis_length_dev, int64_t *sums24_dev, void *scratch_dev, size_t scratch_bytes,
                       void *stream);

/* ---- a4  mask paste of detectron2's detector_postprocess (ref: model/util.py:45-62) -------------------
 * soft_dev (n,M,M) f32 mask-head probabilities, boxes_dev (n,4) f32 [x0,y0,x1,y1] in frame px,
 * out_dev (n,h,w) uint8 {0,1} = bilinear sample (align_corners=False, zeros outside) >= threshold */
MSQ_API int msq_paste_masks(const float *soft_dev, const float *boxes_dev, int n, int M, int h, int w,
                    float threshold, uint8_t *out_dev, void *stream);

/* ---- a8-a10  angles, keypoint flips, iterative angle filter (ref: proc/proc.py:720-724, 827-839,
 *              flips_from_keypoints :851-889, filter_angles :600-624, iterative_filter_angles :627-654)
 * In : orientation_rad (n), axis_length (n,2), centroid (n,2), keypoints (n,8,3) f32 (NaN = no instance)
 * Out: angle_deg (n) f64 final orientation in degrees, flips (n) u8, flip_conf (n) f64 or NULL,
 *      filter_passes (n_chunks) int32 or NULL.
 * The filter is chunk-local: frames [c*chunk, (c+1)*chunk) are filtered independently, which is how the
 * reference behaves when each chunk goes through instances_to_features on its own. */
MSQ_API int msq_angles_and_flips(const double *orientation_rad_dev, const double *axis_length_dev,
                         const double *centroid_dev, const float *keypoints_dev, int n, int chunk,
                         double *angle_deg_dev, uint8_t *flips_dev, double *flip_conf_dev,
                         int32_t *filter_passes_dev, void *stream);

/* the same pieces on their own, for callers that use the reference functions individually:
 * flips_from_keypoints(keypoints, centroids, angles_deg, length) (ref: proc/proc.py:851-889) and
 * iterative_filter_angles(angles, window, tolerance, max_iters) (ref: proc/proc.py:627-654; window <= 15) */
MSQ_API int msq_flips_from_keypoints(const float *keypoints_dev, const double *centroid_dev,
                                     const double *angles_deg_dev, const double *lengths_dev, int n,
                                     uint8_t *flips_dev, double *flip_conf_dev, void *stream);
MSQ_API int msq_iterative_filter_angles(const double *angles_deg_dev, int n, int chunk, int window,
                                        double tolerance, int max_iters, double *out_dev, uint8_t *flips_dev,
                                        int32_t *filter_passes_dev, void *stream);

/* ---- a11 + a12  keypoints_to_dict and compute_scalars (ref: proc/keypoints.py:93-165,
 *                 proc/scalars.py:36-120, proc/util.py:29-61) ---------------------------------------
 * chunk_dev (n,h,w) u8 prepared frames, mask_dev (n,h,w) u8 or NULL (= ones), cleaned_dev (n,h,w) u8.
 * scalars_dev  (17,n) f64 in the order of msq_scalar_name(i)   (area_px exact integer, height_ave_mm
 *              rounded to float32 like the reference's array);
 * kpt_cols_dev (96,n) f64 in the order of msq_keypoint_col_name(i).
 * Either output may be NULL.  scratch_dev: msq_scalars_scratch_bytes(n) bytes, 8-byte aligned.
 * Velocities are first-differences within each `chunk`-frame block (first frame of a block -> 0). */
MSQ_API size_t msq_scalars_scratch_bytes(int n);
MSQ_API int msq_scalars_and_keypoints(const uint8_t *chunk_dev, const uint8_t *mask_dev, const uint8_t *cleaned_dev,
                              const double *centroid_dev, const double *angle_deg_dev,
                              const double *axis_length_dev, const float *keypoints_dev,
                              int n, int h, int w, int chunk, double min_height, double max_height,
                              double true_depth, double *scalars_dev, double *kpt_cols_dev,
                              void *scratch_dev, size_t scratch_bytes, void *stream);
MSQ_API const char *msq_scalar_name(int i);
MSQ_API const char *msq_keypoint_col_name(int i);

/* float64-keypoint variants (the tracking branch writes smoothed float64 keypoints back, proc/proc.py:752) */
MSQ_API int msq_flips_from_keypoints_f64(const double *kpts_dev, const double *centroid_dev, const double *angles_dev,
                                 const double *lengths_dev, int n, uint8_t *flips_dev, double *conf_dev, void *stream);
MSQ_API int msq_scalars_and_keypoints_f64(const uint8_t *chunk_dev, const uint8_t *mask_dev, const uint8_t *cleaned_dev,
                                  const double *centroid_dev, const double *angle_deg_dev, const double *axis_dev,
                                  const double *kpts_dev, int n, int h, int w, int chunk, double min_height,
                                  double max_height, double true_depth, double *scalars_dev, double *kpt_cols_dev,
                                  void *scratch_dev, size_t scratch_bytes, void *stream);

/* ---- a13  crop_and_rotate_frame (ref: proc/proc.py:305-335; pipeline/process_features_step.py:190-195)
 * For every frame i: OpenCV-exact fixed-point bilinear warp of src[i] (h,w) u8 about centroid[i] by
 * angle_deg[i] into out[i] (crop_h,crop_w) u8; NaN / negative centre -> zeros.  src2/out2 optional
 * second plane (the mask) warped with the same transform.  scratch_dev: msq_crop_scratch_bytes(n) bytes,
 * 16-byte aligned (per-frame float64 rotation coefficients + OpenCV's fixed-point row/column tables).
 * Crops up to 512x512; any number of frames per call. */
MSQ_API size_t msq_crop_scratch_bytes(int n);
MSQ_API int msq_crop_rotate(const uint8_t *src_dev, const uint8_t *src2_dev, int n, int h, int w,
                    const double *centroid_dev, const double *angle_deg_dev, int crop_w, int crop_h,
                    uint8_t *out_dev, uint8_t *out2_dev, void *scratch_dev, size_t scratch_bytes, void *stream);

/* ---- a14  Kalman tracking branch ------------------------------------------------------------
 * Replaces pykalman.KalmanFilter as driven by the reference's KalmanTracker (proc/kalman.py:281-418:
 * initialize -> em(n_iter=10) :322-338, smooth_update :386-401, filter_update :408-418, sample :370-377)
 * and the per-frame angle heuristic of instances_to_features (proc/proc.py:771-796).
 * Time-invariant model x' = A x + N(0,Q), z = H x + N(0,R): A (S,S), H (O,S), Q (S,S), R (O,O), m0 (S),
 * P0 (S,S), all float64 row-major on the device; obs (T,O) float64, a row with any non-finite entry is a
 * missing observation (pykalman skips it entirely).  1 <= O <= 32, O <= S <= 64.
 * workspace: msq_kalman_workspace_bytes(T,S,O,for_em) bytes, 256-byte aligned.
 *
 * msq_kalman_smooth: forward filter (the first prediction is (m0,P0) itself unless predict_first != 0,
 *   which is KalmanFilter.filter_update's "predict, then correct") and, when smooth != 0, the RTS smoother.
 *   means_out (T,S) smoothed (or filtered) state means; last_mean (S) / last_cov (S,S) = the filtered state
 *   at T-1, which the reference carries into the next chunk (kalman.py:399-400).  Outputs may be NULL.
 * msq_kalman_em: n_iter EM iterations for transition_covariance, observation_covariance and
 *   initial_state_covariance (the em_vars of kalman.py:326); Q, R, P0 are updated in place.  T >= 2.
 * msq_keypoint_alignment_scores: compute_keypoint_alignment_scores(rotate_points_batch(kpts[:, :7, :2], ...))
 *   (proc/proc.py:762-763, 936-958); kpts (n,8,kp_stride) float64, kp_stride 2 or 3.
 * msq_track_angles: the sequential loop of proc/proc.py:771-796 on the angle tracker (state = (sin,cos) x order,
 *   S = 2*order <= 8): per frame read the tracked angle from `mean`, replace / flip the observed angle
 *   (alignment score < 0.4 -> tracked angle; |difference| > 140 deg -> +180 deg and toggle flips[i]), then
 *   filter_update with it.  angles (n) and flips (n) are updated in place, mean/cov hold the last state. */
MSQ_API size_t msq_kalman_workspace_bytes(int T, int S, int O, int for_em);
MSQ_API int msq_kalman_smooth(const double *A_dev, const double *H_dev, const double *Q_dev, const double *R_dev,
                      const double *m0_dev, const double *P0_dev, const double *obs_dev, int T, int S, int O,
                      int predict_first, int smooth, double *means_out_dev, double *last_mean_dev,
                      double *last_cov_dev, void *workspace_dev, size_t workspace_bytes, void *stream);
MSQ_API int msq_kalman_em(const double *A_dev, const double *H_dev, double *Q_dev, double *R_dev, const double *m0_dev,
                  double *P0_dev, const double *obs_dev, int T, int S, int O, int n_iter, void *workspace_dev,
                  size_t workspace_bytes, void *stream);
MSQ_API int msq_keypoint_alignment_scores(const double *kpts_dev, int kp_stride, const double *centroid_dev,
                                  const double *angles_deg_dev, int n, double *scores_dev, void *stream);
/* a8 + a9 of the tracking branch (proc/proc.py:720-724, 756-763): orientation (rad) -> clamped degrees, keypoint flip
 * votes on float64 (smoothed) keypoints (n,8,3), angles[flips] = clamp(angles + 180), alignment scores of the
 * keypoints rotated by the result.  conf may be NULL. */
MSQ_API int msq_tracking_prepare(const double *orientation_rad_dev, const double *axis_length_dev, const double *centroid_dev,
                         const double *kpts_dev, int n, double *angles_deg_dev, uint8_t *flips_dev, double *conf_dev,
                         double *scores_dev, void *stream);
MSQ_API int msq_track_angles(const double *A_dev, const double *H_dev, const double *Q_dev, const double *R_dev,
                     double *mean_dev, double *cov_dev, int S, double *angles_deg_dev, uint8_t *flips_dev,
                     const double *scores_dev, int n, void *stream);

/* ---- f4  get_bground_im (session setup; ref: proc/roi.py:293-307, io/session.py:217-218) --------------
 * frames_dev (n,H,W) 16-bit (int16 as read by read_frames_raw, or uint16 with is_unsigned != 0): every frame is
 * median-blurred with a med_scale x med_scale window (3 or 5, replicated border = cv2.medianBlur) into scratch --
 * the input is NOT modified, unlike the reference -- and out_dev (H,W) float64 receives the per-pixel median over
 * the n blurred frames (np.median: mean of the two middle values for even n).  Bit-exact.
 * scratch_dev: msq_bground_scratch_bytes(n,H,W) bytes.  n <= 3200. */
MSQ_API size_t msq_bground_scratch_bytes(int n, int H, int W);
MSQ_API int msq_get_bground_im(const void *frames_dev, int n, int H, int W, int med_scale, int is_unsigned,
                       double *out_dev, void *scratch_dev, size_t scratch_bytes, void *stream);

/* ---- f4  get_roi (session setup; ref: proc/roi.py:14-103 get_roi, :106-130 plane_fit3, :133-212 plane_ransac) ------
 * The host side draws the RANSAC triples (np.random, the reference's order) and ranks the regions; these entry points do
 * the per-pixel work.  depth_dev is the (H,W) background image as float64.
 *
 * msq_sobel_gradient_mask (gradient_filter, proc/roi.py:29-35): mask_dev (H*W) u8 = |cv2.Sobel(depth, CV_64F, 1, 0, k)| <
 *   threshold and the same for (0, 1); deriv_host / smooth_host are cv2.getDerivKernels' taps in HOST memory (odd counts
 *   <= 31), border = BORDER_REFLECT_101.  Exact for integer- or half-integer-valued images (all sums are exact in float64).
 * msq_plane_ransac_score: idx_dev (npts) int32 = raster indices of the usable pixels (depth_range / gradient mask);
 *   sel_dev (iters,3) int64 indexes idx_dev.  Per candidate: planes_dev (iters,4) = unit normal and offset of the plane
 *   through its 3 points (NaN when they are collinear), ninliers_dev = #{|ax+by+cz+d| < tol}, sumdist_dev = sum of the
 *   distances over the usable pixels (the reference's np.mean(dist) times npts; summation order differs).
 * msq_plane_distance: dist_dev (H*W) float64 and/or on_plane_dev (H*W) u8 = dist < tol (and valid_dev != 0 when given);
 *   plane_host = 4 doubles in HOST memory.
 * msq_label_regions: 8-connected labels of on-plane pixels, 0 = background, regions numbered 1.. in raster order of
 *   their first pixel (skimage.measure.label); n_regions_dev receives the count.  Bit-exact.
 * msq_region_props: per region r (label r+1): area, bbox (ymin, xmin, ymax, xmax inclusive) and maxd4 = max over its
 *   pixels of (2y-H)^2 + (2x-W)^2, i.e. 4x the squared distance to the image centre the reference ranks by.
 * msq_region_rois: for the regions listed in order_dev (0-based region ids): cv2.dilate by se_dilate (dh x dw u8,
 *   anchored at its centre; NULL = skip), cv2.erode by se_erode (NULL = skip), scipy binary_fill_holes when
 *   fill_holes != 0; rois_dev (n_out,H,W) u8 0/1, bboxes_dev (n_out,4) = get_bbox of each mask (-1 when empty).
 *   W <= 1024, H * ceil(W/32) * 8 bytes of shared memory <= 220 KB.  Bit-exact. */
MSQ_API int msq_sobel_gradient_mask(const double *depth_dev, int H, int W, const double *deriv_host, int n_deriv,
                            const double *smooth_host, int n_smooth, double threshold, uint8_t *mask_dev, void *stream);
MSQ_API int msq_plane_ransac_score(const int *idx_dev, int npts, const double *depth_dev, int H, int W, const long long *sel_dev,
                           int iters, double tol, double *planes_dev, int *ninliers_dev, double *sumdist_dev, void *stream);
MSQ_API int msq_plane_distance(const double *depth_dev, int H, int W, const double *plane_host, double tol,
                       const uint8_t *valid_dev, double *dist_dev, uint8_t *on_plane_dev, void *stream);
MSQ_API size_t msq_label_scratch_bytes(int H, int W);
MSQ_API int msq_label_regions(const uint8_t *bin_dev, int H, int W, int *labels_dev, int *n_regions_dev, void *scratch_dev,
                      size_t scratch_bytes, void *stream);
MSQ_API int msq_region_props(const int *labels_dev, int H, int W, int n_regions, int *area_dev, int *bbox_dev,
                     unsigned *maxd4_dev, void *stream);
MSQ_API int msq_region_rois(const int *labels_dev, int H, int W, const int *order_dev, int n_out, const uint8_t *se_dilate_dev,
                    int dh, int dw, const uint8_t *se_erode_dev, int eh, int ew, int fill_holes, uint8_t *rois_dev,
                    int *bboxes_dev, void *stream);

/* ---- whole-chunk pipeline: everything ProcessFeaturesStep.process does (ref:
 *      pipeline/process_features_step.py:56-60,163-199 with use_tracking=False), device buffers -------
 * Everything is ordered on `stream`; internally the few frames the streaming feature kernel hands to the general one run
 * on a library-owned side stream that is forked from and joined back into `stream` with events (no host
 * synchronisation, capturable in a CUDA graph).  scratch_dev: msq_extract_scratch_bytes(n,h,w) bytes, 256-byte aligned. */
typedef struct msq_chunk_outputs {
    uint8_t *cleaned;        /* (n,h,w)  */
    double  *centroid;       /* (n,2)    */
    double  *angle_deg;      /* (n)      */
    double  *axis_length;    /* (n,2)    */
    uint8_t *flips;          /* (n)      */
    double  *scalars;        /* (17,n)   */
    double  *kpt_cols;       /* (96,n)   */
    uint8_t *depth_crops;    /* (n,crop_h,crop_w) */
    uint8_t *mask_crops;     /* (n,crop_h,crop_w) */
    int32_t *filter_passes;  /* (ceil(n/chunk)) or NULL */
} msq_chunk_outputs;

MSQ_API size_t msq_extract_scratch_bytes(int n, int h, int w);
/* The side stream + fork / join events msq_extract_chunk runs its masked sums on (beside the short launches of the main
 * stream), as an explicit object of the CURRENT device: create once per (thread, GPU), pass to msq_extract_chunk_engine,
 * destroy at the end.  msq_extract_chunk is the same call with an engine the library keeps per (host thread, device) -- the
 * only state the library ever holds -- and without positive-pixel rows.  positive_bits_dev: see msq_prep_frames_bits, or NULL. */
typedef struct msq_engine msq_engine;
MSQ_API int msq_engine_create(msq_engine **engine);
MSQ_API int msq_engine_destroy(msq_engine *engine);
MSQ_API int msq_extract_chunk_engine(msq_engine *engine, const uint8_t *chunk_dev, const uint32_t *positive_bits_dev,
                      const uint8_t *mask_dev, const float *keypoints_dev, int n, int h, int w, int chunk, double min_height, double max_height, double true_depth,
                      int crop_w, int crop_h, const msq_chunk_outputs *out, void *scratch_dev,
                      size_t scratch_bytes, void *stream);
MSQ_API int msq_extract_chunk(const uint8_t *chunk_dev, const uint8_t *mask_dev, const float *keypoints_dev,
                      int n, int h, int w, int chunk, double min_height, double max_height, double true_depth,
                      int crop_w, int crop_h, const msq_chunk_outputs *out, void *scratch_dev,
                      size_t scratch_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MOSEQ_B200_H */
