/*
 * ssdx.h - C ABI of libssdx.so: the B200 (sm_100a) SSD target/prediction path.
 *
 * The reference (matteo-stat/multi-task-learning-object-detection-semantic-
 * segmentation, package `ssdseglib`) is pure Python and has no FFI of its own;
 * its boundary for this path is the Python API of ssdseglib/boxes.py,
 * datacoder.py, layers.py and evaluators.py.  Each entry point below replaces
 * the arithmetic behind one of those Python entry points (cited as
 * file:line); the Python mirror in
 * multi-task-learning-object-detection-semantic-segmentation_b200/ssdseglib/
 * binds them through ctypes (INTEGRATION.md shows the stubs).
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; ssdx_last_error()
 *     returns a thread-local message for the last failure on this thread;
 *   - unless a parameter is documented "host", data pointers are DEVICE
 *     pointers owned by the caller and living on the device the anchor handle
 *     (or the `device` argument) names; all arrays are dense, row-major;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default
 *     stream).  Calls are asynchronous, stream-ordered, allocate nothing and
 *     never synchronise; per-call scratch lives in the caller's workspace;
 *   - all exports are re-entrant (the reference's encode is called from
 *     concurrent tf.data workers, 03-...ipynb cell 8); the anchor handle is
 *     immutable after creation and may be shared across threads and streams;
 *   - there is no CPU fallback: without a CUDA device every compute entry
 *     point fails with SSDX_ERR_CUDA.
 */
#ifndef SSDX_H_
#define SSDX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSDX_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define SSDX_API __attribute__((visibility("default")))
#else
#define SSDX_API
#endif

#define SSDX_OK 0
#define SSDX_ERR_INVALID (-1)   /* bad argument (mirrors the reference's ValueError sites) */
#define SSDX_ERR_CUDA (-2)      /* CUDA runtime / launch failure */
#define SSDX_ERR_WORKSPACE (-3) /* caller workspace too small */
#define SSDX_ERR_NCCL (-4)      /* NCCL missing in the process, or a collective / communicator call failed */

typedef struct ssdx_anchors ssdx_anchors; /* opaque, device-resident anchor table */
typedef void* ssdx_stream;                /* cudaStream_t */

SSDX_API int ssdx_version(void);
SSDX_API const char* ssdx_last_error(void);
/* Number of CUDA devices visible; SSDX_ERR_CUDA if the driver is unusable. */
SSDX_API int ssdx_device_count(int* count);
/* Process-level dispatch override; every path of an entry point gives identical results, the options exist for tests
 * and tuning runs.  key "encode_path": "auto" | "grid_fused" | "grid_split" (closed-form windows, anchor tables that
 * are unions of regular grids) | "tile_fused" | "tile_tiled" (tile match, any anchor table);
 * key "nms_path": "auto" | "perclass" | "pcta" | "generic" | "blocked"; key "nms_cta": "auto" | "256" | "512" | "1024"
 * (threads per image in the merged NMS pipeline; auto = from the batch size).  A forced path that cannot serve a call
 * (table is not a grid, shapes too large for one CTA per image) falls back to the automatic choice for that call. */
SSDX_API int ssdx_set_option(const char* key, const char* value);

/* ---- anchors: datacoder.py:6-117 (ctor), boxes.py:391-445 (conversions) ----
 * corners_xyxy_host : host, (A,4) xmin,ymin,xmax,ymax - what the ctor stores or derives (:59-62, :84-89)
 * centroids_host    : host, (A,4) cx,cy,w,h used by the DECODE side (:65-70, :78-81); NULL = derive from corners.
 * The encode side always re-derives centroids from the corners (datacoder.py:251-256); the area is
 * (ymax-ymin+1)*(xmax-xmin+1) (:111-114).  The table is uploaded once and stays in HBM/L2. */
SSDX_API int ssdx_anchors_create(const float* corners_xyxy_host, const float* centroids_host, int32_t num_anchors, int device,
                        ssdx_anchors** out);
SSDX_API int ssdx_anchors_destroy(ssdx_anchors* anchors);
SSDX_API int ssdx_anchors_count(const ssdx_anchors* anchors, int32_t* num_anchors);
SSDX_API int ssdx_anchors_device(const ssdx_anchors* anchors, int* device);
/* num_grids > 0 when the table was recognised as a union of regular (feature map, box shape) grids - true for every
 * output of boxes.DefaultBoundingBoxes (boxes.py:112-146, :198-205); encode then finds the pairs that can matter in
 * closed form.  num_spans = row ranges the large-table / small-batch encode driver cuts the maps into.  Either
 * pointer may be NULL. */
SSDX_API int ssdx_anchors_grid_info(const ssdx_anchors* anchors, int32_t* num_grids, int32_t* num_spans);
/* The same detection on the host alone (no device): grids_out, when not NULL, receives up to `capacity` rows of
 * (first anchor, rows, columns, shapes per cell) - anchor(row, col) of grid i = first + (row * columns + col) * shapes. */
SSDX_API int ssdx_grid_probe(const float* corners_xyxy_host, int32_t num_anchors, int32_t* num_grids,
                             int32_t* num_spans, int32_t* grids_out, int32_t capacity);
/* Host-only: the cell window ssdx_encode's grid drivers enumerate for every (GT box i, grid q) at the IoU bound
 * t_host[i] - no anchor outside the window reaches IoU >= t with the box (datacoder.py:205-220 arithmetic).
 * out (n_boxes, num_grids, 4) i32 = (first row, rows, first column, columns); rows = 0: empty window. */
SSDX_API int ssdx_grid_window_probe(const float* corners_xyxy_host, int32_t num_anchors, const float* gt_xyxy_host,
                                    const float* t_host, int32_t n_boxes, int32_t* out);

/* ---- encode: datacoder.py:205-300 (_encode_ground_truth_labels_boxes, after the CSV read) ----
 * gt_xyxy   (B,Gmax,4) f32, gt_label (B,Gmax) i32, gt_count (B) i32 (0 <= count <= Gmax; 0 => all background)
 * std_host  : host, 4 floats (sigma cx, cy, w, h), datacoder.py:51
 * out_labels (B,A,C) f32 one-hot (background row = [1,0,...]), out_offsets (B,A,4) f32 (tx,ty,tw,th)
 * out_match  (B,A) i32 matched GT index or -1; may be NULL
 * IoU, both arg-max reductions (ties -> lowest index) and the duplicate-resolution rule are bit-exact
 * restatements of :205-244 and of the sequential scatter :286-298 (DESIGN.md "Parity contract"). */
SSDX_API size_t ssdx_encode_workspace_bytes(int32_t num_anchors, int32_t batch, int32_t gmax);
SSDX_API int ssdx_encode(const ssdx_anchors* anchors, const float* gt_xyxy, const int32_t* gt_label, const int32_t* gt_count,
                int32_t batch, int32_t gmax, int32_t num_classes, float iou_threshold, const float* std_host,
                float* out_labels, float* out_offsets, int32_t* out_match, void* workspace, size_t workspace_bytes,
                ssdx_stream stream);

/* ---- decode: layers.py:45-81, datacoder.py:349-388, :390-432 ---- */
#define SSDX_DECODE_YXYX 0             /* layers.DecodeBoxesCentroidsOffsets.call: (ymin,xmin,ymax,xmax), unmasked */
#define SSDX_DECODE_CENTROIDS_MASKED 1 /* DataEncoderDecoder.decode_to_centroids: (cx,cy,w,h), background rows 0 */
#define SSDX_DECODE_CORNERS_MASKED 2   /* DataEncoderDecoder.decode_to_corners: (xmin,ymin,xmax,ymax), masked twice */
SSDX_API int ssdx_decode(const ssdx_anchors* anchors, const float* offsets /*(B,A,4)*/, int32_t batch, const float* std_host,
                int mode, float* out /*(B,A,4)*/, ssdx_stream stream);

/* ---- NMS: layers.py:127-168 -> tf.image.combined_non_max_suppression(q=1, clip_boxes=False) ----
 * boxes_yxyx (B,A,4), scores (B,A,C) (background column included, as in the reference)
 * out (B,K,6) = [label, prob, xmin, ymin, xmax, ymax] (layers.py:155-162), K = max_total; unused rows are 0
 * out_valid (B) i32 = rows in use (TF's valid_detections, discarded by the layer); may be NULL
 * out_anchor_idx (B,K) i32 = source anchor of each row, -1 padding (keep-set parity); may be NULL
 * Equal scores are visited lower anchor index first, and merged lower class first (TF leaves both unspecified). */
SSDX_API size_t ssdx_nms_workspace_bytes(int32_t batch, int32_t num_anchors, int32_t num_classes, int32_t max_per_class,
                                int32_t max_total);
SSDX_API int ssdx_nms(const float* boxes_yxyx, const float* scores, int32_t batch, int32_t num_anchors, int32_t num_classes,
             int32_t max_per_class, int32_t max_total, float iou_threshold, float score_threshold, float* out,
             int32_t* out_valid, int32_t* out_anchor_idx, void* workspace, size_t workspace_bytes, int device,
             ssdx_stream stream);

/* Same, with every score multiplied by class_scale[c] (device, C floats; NULL = 1) as it is read: the product
 * layers.SegmentationSuppression feeds to the NMS layer (layers.py:209, models.py:402-414) without materialising the
 * suppressed (B,A,C) tensor.  One IEEE multiply per score, so results equal NMS over the multiplied tensor. */
SSDX_API int ssdx_nms_scaled(const float* boxes_yxyx, const float* scores, const float* class_scale, int32_t batch,
                             int32_t num_anchors, int32_t num_classes, int32_t max_per_class, int32_t max_total,
                             float iou_threshold, float score_threshold, float* out, int32_t* out_valid,
                             int32_t* out_anchor_idx, void* workspace, size_t workspace_bytes, int device,
                             ssdx_stream stream);

/* The same NMS in two stream-ordered phases on one workspace.  ssdx_nms_select reads only the scores (candidate lists
 * into the workspace); ssdx_nms_finish needs the boxes and must see the same scores, class_scale, shape, caps,
 * score_threshold and workspace.  In the reference's inference graph the two inputs of the NMS layer are independent
 * branches (models.py:402-414: output-labels -> [SegmentationSuppression], output-boxes -> Decode), so phase 1 may
 * run on another stream while the boxes are still being decoded; the caller orders the two calls (events).
 * select + finish == ssdx_nms_scaled, bit for bit. */
SSDX_API int ssdx_nms_select(const float* scores, const float* class_scale, int32_t batch, int32_t num_anchors,
                             int32_t num_classes, int32_t max_per_class, int32_t max_total, float score_threshold,
                             void* workspace, size_t workspace_bytes, int device, ssdx_stream stream);
SSDX_API int ssdx_nms_finish(const float* boxes_yxyx, const float* scores, const float* class_scale, int32_t batch,
                             int32_t num_anchors, int32_t num_classes, int32_t max_per_class, int32_t max_total,
                             float iou_threshold, float score_threshold, float* out, int32_t* out_valid,
                             int32_t* out_anchor_idx, void* workspace, size_t workspace_bytes, int device,
                             ssdx_stream stream);

/* Decode fused into the NMS: the wiring of get_model_for_inference (models.py:410-414: DecodeBoxesCentroidsOffsets ->
 * NonMaximumSuppression) whose only output is the detections.  offsets (B,A,4) are the raw regression outputs; every
 * box the NMS touches (a few hundred candidates per image) is decoded on the fly with the decode kernel's own
 * formula, so the result equals ssdx_decode(mode YXYX) followed by ssdx_nms_scaled bit for bit, without writing and
 * re-reading the (B,A,4) box tensor.  A and the device come from the anchor handle; std_host as in ssdx_decode. */
SSDX_API int ssdx_decode_nms(const ssdx_anchors* anchors, const float* offsets, const float* std_host,
                             const float* scores, const float* class_scale, int32_t batch, int32_t num_classes,
                             int32_t max_per_class, int32_t max_total, float iou_threshold, float score_threshold,
                             float* out, int32_t* out_valid, int32_t* out_anchor_idx, void* workspace,
                             size_t workspace_bytes, ssdx_stream stream);

/* ---- segmentation suppression: layers.py:189-212 (SegmentationSuppression.call), SURVEY 8(f) row 2 ----
 * seg_probs (num_pixels, Cm) f32 = the (B,H,W,Cm) segmentation output flattened over batch and space.
 * out_presence (depth) f32: 1.0f iff some pixel of the WHOLE batch has argmax == c (ties -> lowest index), else 0.0f
 * (argmax -> one_hot(depth) -> reduce_sum over (0,1,2) -> clip to [0,1]; the reference hard-codes depth = 4).
 * workspace: 256 bytes, 256-byte aligned. */
SSDX_API int ssdx_seg_class_presence(const float* seg_probs, int64_t num_pixels, int32_t cm, int32_t depth,
                                     float* out_presence, void* workspace, size_t workspace_bytes, int device,
                                     ssdx_stream stream);
/* out[r][c] = scores[r][c] * class_scale[c] (layers.py:209 tf.math.multiply with broadcasting); rows = B*A. */
SSDX_API int ssdx_scale_scores(const float* scores, int64_t rows, int32_t num_classes, const float* class_scale,
                               float* out, int device, ssdx_stream stream);

/* ---- detection eval: evaluators.py:6-63 (_iou_boxes_pred_vs_true) + :106-157 (TP flags, GT counts) ----
 * pred_label (B,K) i32, pred_xyxy (B,K,4) f32; gt_* as in encode; labels are class indices in [0,C)
 * out_tp (B,K) u8: 1 iff label != background and max_g IoU*[label match] >= iou_threshold (fp32 compare)
 * counts (C,3) i64 ACCUMULATED (caller zeroes): [c][0] += TP, [c][1] += FP, [c][2] += GT boxes of class c */
SSDX_API int ssdx_eval_tp(const int32_t* pred_label, const float* pred_xyxy, const float* gt_xyxy, const int32_t* gt_label,
                 const int32_t* gt_count, int32_t batch, int32_t k, int32_t gmax, int32_t num_classes,
                 int32_t background_label, float iou_threshold, uint8_t* out_tp, int64_t* counts, int device,
                 ssdx_stream stream);

/* ---- the one collective of the path: all-gather of the per-rank (C,3) count tables (BASELINE config 5) ----
 * Evaluation shards the images over the ranks of one box (evaluators.py:106-157 per shard, no communication); the
 * (TP, FP, n_gt) tables are then gathered over NCCL so that every rank can form the global counts that
 * evaluators.py:159-187 integrates.  libssdx binds to the NCCL already loaded in the process (libnccl.so.2, or
 * $SSDX_NCCL_LIB) at first use - no link-time dependency; without NCCL these four calls return SSDX_ERR_NCCL.
 *   ssdx_comm_unique_id : rank 0 draws the 128-byte NCCL unique id and hands it to the other ranks out of band
 *                         (the Python mirror broadcasts it with torch.distributed);
 *   ssdx_comm_create    : collective over the nranks ranks (ncclCommInitRank), one communicator per (process, device);
 *   ssdx_allgather_counts: counts (C,3) i64 device -> gathered (nranks,C,3) i64 device, rank-major; stream-ordered
 *                         (ncclAllGather on `stream`), every rank of the communicator must call it. */
typedef struct ssdx_comm ssdx_comm;
SSDX_API int ssdx_nccl_version(int* version);
SSDX_API int ssdx_comm_unique_id(uint8_t* id128);
SSDX_API int ssdx_comm_create(const uint8_t* id128, int32_t nranks, int32_t rank, int device, ssdx_comm** out);
SSDX_API int ssdx_comm_destroy(ssdx_comm* comm);
SSDX_API int ssdx_allgather_counts(ssdx_comm* comm, const int64_t* counts, int64_t* gathered, int32_t num_classes,
                                   ssdx_stream stream);
/* Everything a rank needs for the global AP tail (evaluators.py:159-187) in ONE NCCL launch (ncclGroupStart/End): the
 * count tables as above plus the TP flags, confidences and labels of every rank's n_local predictions (the same n_local
 * on every rank), rank-major = image-major for contiguous shards: all_tp / all_confidence / all_label (nranks * n_local). */
SSDX_API int ssdx_allgather_eval(ssdx_comm* comm, const int64_t* counts, int64_t* gathered_counts, int32_t num_classes,
                                 const uint8_t* tp, const float* confidence, const int32_t* label, int64_t n_local,
                                 uint8_t* all_tp, float* all_confidence, int32_t* all_label, ssdx_stream stream);
/* out (C,3) i64 = sum over the nranks tables of gathered (nranks,C,3): the global counts on every rank. */
SSDX_API int ssdx_sum_counts(const int64_t* gathered, int32_t nranks, int32_t num_classes, int64_t* out, int device,
                             ssdx_stream stream);

/* ---- average precision: evaluators.py:159-187 (the tail of average_precision_object_detection), SURVEY 8(f) row 4 ----
 * tp (N) u8, confidence (N) f32, label (N) i32 class indices: every prediction of the data set, flattened;
 * counts (C,3) i64 as ssdx_eval_tp accumulates them ([c][2] = ground-truth boxes of class c).
 * out_ap (C) f64: per class, predictions sorted by confidence descending, precision = cumsum(tp)/rank (f64),
 * recall = cumsum(tp)/n_gt (f32), trapezoid in f64 without a (0,.) sentinel; 0 for the background class, for a class
 * without ground truth and for a class without predictions.  N < 2^31.  Tie order among equal confidences is the
 * input order (the reference's quicksort leaves it unspecified). */
SSDX_API size_t ssdx_average_precision_workspace_bytes(int64_t num_predictions, int32_t num_classes);
SSDX_API int ssdx_average_precision(const uint8_t* tp, const float* confidence, const int32_t* label,
                                    int64_t num_predictions, int32_t num_classes, const int64_t* counts,
                                    int32_t background_label, double* out_ap, void* workspace, size_t workspace_bytes,
                                    int device, ssdx_stream stream);

/* evaluators.py:6-63 as a matrix: out (P,G) f32 = IoU (+1 convention, +1e-7f in the denominator) * [labels equal].
 * G == 0 is the caller's business (the reference returns a (P,1) zero matrix, :27-28). */
SSDX_API int ssdx_iou_pred_vs_true(const int32_t* pred_label, const float* pred_xyxy, const int32_t* gt_label,
                                   const float* gt_xyxy, int32_t num_pred, int32_t num_gt, float* out, int device,
                                   ssdx_stream stream);

/* ---- class head: soft-max of the raw logits, and the NMS taking logits directly ----
 * The reference's class head ends in tf.keras.layers.Softmax (models.py:259, :694) whose (B,A,C) output feeds
 * SegmentationSuppression and the NMS layer.  ssdx_softmax_scores writes probabilities = softmax(logits) [* class_scale]
 * (rows = B*A; shifted by the row maximum, hardware exponential, fp32 sum in class order, one hardware reciprocal: within
 * ~4e-6 relative of the exact soft-max, inside the path's 1e-5 fp32 tolerance).  ssdx_nms_logits == ssdx_softmax_scores
 * followed by ssdx_nms_scaled, bit for bit, but with 4 classes (rows of 16 aligned bytes) the probabilities are
 * computed inside the kernels that read the scores and never written to memory; otherwise they are materialised in the
 * workspace first. */
SSDX_API int ssdx_softmax_scores(const float* logits, int64_t rows, int32_t num_classes, const float* class_scale,
                                 float* out, int device, ssdx_stream stream);
SSDX_API size_t ssdx_nms_logits_workspace_bytes(int32_t batch, int32_t num_anchors, int32_t num_classes,
                                                int32_t max_per_class, int32_t max_total);
SSDX_API int ssdx_nms_logits(const float* boxes_yxyx, const float* logits, const float* class_scale, int32_t batch,
                             int32_t num_anchors, int32_t num_classes, int32_t max_per_class, int32_t max_total,
                             float iou_threshold, float score_threshold, float* out, int32_t* out_valid,
                             int32_t* out_anchor_idx, void* workspace, size_t workspace_bytes, int device,
                             ssdx_stream stream);

/* The whole inference tail of get_model_for_inference (models.py:402-414) in one call: == ssdx_decode(mode YXYX) +
 * ssdx_softmax_scores + ssdx_nms_scaled, bit for bit, with the boxes decoded per candidate and the probabilities
 * computed inside the score filter (4 classes, max_per_class >= max_total; otherwise the probabilities are materialised
 * in the workspace).  Workspace: ssdx_nms_logits_workspace_bytes. */
SSDX_API int ssdx_decode_nms_logits(const ssdx_anchors* anchors, const float* offsets, const float* std_host,
                                    const float* logits, const float* class_scale, int32_t batch, int32_t num_classes,
                                    int32_t max_per_class, int32_t max_total, float iou_threshold, float score_threshold,
                                    float* out, int32_t* out_valid, int32_t* out_anchor_idx, void* workspace,
                                    size_t workspace_bytes, ssdx_stream stream);

/* ---- training-time box metric: metrics.py:53-173 (jaccard_iou_bounding_boxes), SURVEY 8(f) row 3 ----
 * y_true, y_pred (B,A,4) f32 standardised centroid offsets (y_true = encode output, background rows all zero);
 * out (B) f32 = mean over the non-background anchors of IoU(decoded prediction, decoded target) with the reference's
 * arithmetic (width/height clamped at 0, +1 pixel convention, epsilon 1e-7); NaN for an image without positives,
 * as the reference (0 / 0).  Uses the decode-side centroids of the anchor handle; std_host as in ssdx_decode. */
SSDX_API int ssdx_metric_iou_boxes(const ssdx_anchors* anchors, const float* y_true, const float* y_pred, int32_t batch,
                                   const float* std_host, float* out, ssdx_stream stream);

/* ---- training losses on the encoded targets: losses.py:5-50 (localization_loss), :52-172 (confidence_loss),
 *      SURVEY 8(f) row 3 ----
 * ssdx_localization_loss: y_true, y_pred (B,A,4) f32 (16-byte aligned); out_loss (B) f32 = smooth-L1 summed over the
 * anchors whose target row is not background (sum |y_true| > 0) / max(their count, 1).  out_grad (B,A,4), nullable:
 * d out_loss[b] / d y_pred[b] as TensorFlow's autodiff derives it (-(d) inside |d| < 1, -sign(d) outside, 0 on
 * background rows, all over the count).
 * ssdx_confidence_loss: y_true (B,A,C) one-hot targets (column 0 = background indicator), y_pred (B,A,C)
 * probabilities; clip_min / clip_max = the fp32 values of epsilon and 1 - epsilon (Keras epsilon 1e-7, :89-90).
 * out_loss (B) f32 = (cross-entropy of the non-background anchors + cross-entropy of the kept background anchors) /
 * max(non-background anchors of the image, 1); kept = the k = min(3 * positives, negatives) largest background losses
 * of the WHOLE batch (:117, :134-138), equal losses lower flat index first.  out_grad (B,A,C) nullable: d out_loss[b] /
 * d y_pred[b] (0 where the clip saturates); out_keep (B,A) u8 nullable: the kept background anchors.  Stream-ordered,
 * no host synchronisation (k and the threshold stay on the device). */
SSDX_API int ssdx_localization_loss(const float* y_true, const float* y_pred, int32_t batch, int32_t num_anchors,
                                    float* out_loss, float* out_grad, int device, ssdx_stream stream);
SSDX_API size_t ssdx_confidence_loss_workspace_bytes(int32_t batch, int32_t num_anchors);
SSDX_API int ssdx_confidence_loss(const float* y_true, const float* y_pred, int32_t batch, int32_t num_anchors,
                                  int32_t num_classes, float clip_min, float clip_max, float* out_loss, float* out_grad,
                                  uint8_t* out_keep, void* workspace, size_t workspace_bytes, int device,
                                  ssdx_stream stream);

/* ---- GT ingest (host): datacoder.py:194-203, evaluators.py:111-121 ----
 * Parses `label,xmin,ymin,xmax,ymax` records from `text` (host).  crlf_only != 0 reproduces datacoder.py:195
 * (records split on "\r\n" only); 0 accepts "\n", "\r\n" (csv.reader).  flip_width > 0 applies the horizontal flip
 * xmin' = W - xmax, xmax' = W - xmin in fp32 (:203).  Writes at most `capacity` rows to host arrays. */
SSDX_API int ssdx_parse_gt_csv(const char* text, size_t text_len, int crlf_only, float flip_width, int32_t* labels_host,
                      float* boxes_xyxy_host, int32_t capacity, int32_t* count);

/* Batched form of the ingest (host, `threads` worker threads; <= 0 = all hardware threads): one file per image
 * (datacoder.py:194 tf.io.read_file per element of the tf.data pipeline; evaluators.py:109-121 open() per image).
 * paths (n_files) host C strings; flip (n_files) host bytes or NULL: files with a non-zero byte get the horizontal
 * flip with flip_width (datacoder.py:337-343 draws one coin per image).  Outputs are host arrays, padded with
 * zeros: labels (n_files, gmax) i32, boxes (n_files, gmax, 4) f32, counts (n_files) i32.  If a file holds more than
 * gmax records the call returns SSDX_ERR_WORKSPACE with counts holding the true numbers (retry with a larger gmax;
 * gmax = 0 with NULL arrays only counts). */
SSDX_API int ssdx_read_gt_csv_batch(const char* const* paths, int32_t n_files, int crlf_only, const uint8_t* flip,
                                    float flip_width, int32_t gmax, int32_t* labels_host, float* boxes_xyxy_host,
                                    int32_t* counts_host, int32_t threads);

#ifdef __cplusplus
}
#endif
#endif /* SSDX_H_ */
