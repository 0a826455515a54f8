/*
 * ronk.h -- C ABI of the B200-native detection hot path (libronk.so).
 *
 * Drop-in boundary for HiKapok/RON_Tensorflow's data-parallel hot path.  The
 * reference has no FFI: its boundary is Python functions that build TF-1 graph
 * nodes.  Each entry point below names the reference
 * interface it replaces (file:line in the reference repo).  The reference-named
 * Python modules in ron_tensorflow_b200/{nets,tf_extended} bind these through
 * ctypes (ron_tensorflow_b200/_ffi.py); INTEGRATION.md shows the stub.
 *
 * Conventions
 *  - plain pointers and sizes only; every data pointer is a DEVICE pointer unless the
 *    name ends in _host; `stream` is a cudaStream_t passed as void* (NULL = default);
 *  - every call is asynchronous on `stream`, never synchronises the device, allocates
 *    nothing (scratch comes in through *_workspace_bytes + a caller buffer);
 *  - return 0 on success, <0 = RONK_E*; ronk_last_error() gives a thread-local message;
 *  - boxes are float32 (ymin, xmin, ymax, xmax); localisations are float32 (x, y, w, h);
 *  - anchors are flattened layer-major, then (row, col, a) with a fastest (N total);
 *  - there is NO CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef RONK_H_
#define RONK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RONK_VERSION 100

#define RONK_OK 0
#define RONK_EINVAL (-1)   /* bad argument (NULL, size, alignment)            */
#define RONK_ECUDA (-2)    /* CUDA runtime error (message has the CUDA text)  */
#define RONK_ENOMEM (-3)   /* allocation failed (anchor handle only)          */
#define RONK_ELIMIT (-4)   /* size outside what the kernels support           */

#define RONK_KIND_RON 0    /* nets/ron_vgg_320.py:285-333 anchor rule */
#define RONK_KIND_SSD 1    /* nets/ssd_vgg_512.py:286-338 anchor rule */

#define RONK_NMS_MIN 0     /* overlap = inter / min(area_j, area_i)  tf_extended/bboxes.py:207-208 (default) */
#define RONK_NMS_UNION 1   /* overlap = inter / ((area_j - inter) + area_i)  tf_extended/bboxes.py:205-206 */

/* match flags: the two arguments of do_dual_max_match no caller changes (nets/ssd_common.py:49) */
#define RONK_MATCH_DEFAULT 0
#define RONK_MATCH_NO_IGNORE_BETWEEN 1   /* ignore_between=False */
#define RONK_MATCH_NO_GT_MAX_FIRST 2     /* gt_max_first=False   */

/* select flags */
#define RONK_SELECT_DEFAULT 0
#define RONK_SELECT_LOC_DECODED 1        /* loc_layers already hold decoded boxes (the reference's
                                            detected_bboxes receives bboxes_decode's output) */
#define RONK_SELECT_NO_SAMPLING 2        /* one scatter pass with select_threshold only (no sampled pivot) */
#define RONK_SELECT_TEST_REBUILD 4       /* test hook: pivot = highest sampled score and every pivoted
                                            (image, class) takes the exact list-rebuild path */

typedef struct ronk_anchors ronk_anchors_t;

int ronk_version(void);
/* id of the CUDA-graph capture `stream` takes part in, 0 when it is not capturing (host-side helper of the bindings) */
int ronk_stream_capture_id(void* stream, unsigned long long* out_id);
const char* ronk_last_error(void);

/* ------------------------------------------------------------------ anchors
 * Replaces RONNet.anchors -> ron_anchors_all_layers -> ron_anchor_one_layer
 * (nets/ron_vgg_320.py:162-171,336-355,285-333), ssd_anchors_all_layers
 * (nets/ssd_vgg_512.py:341-358, nets/ssd_vgg_300.py:361-380) and the anchor
 * bookkeeping at the top of tf_ssd_bboxes_encode (nets/ssd_common.py:371-402).
 * Runs the anchor-generator kernel once on the current device and keeps, in HBM:
 *   table 0  decode anchors   float32[N,4] (y, x, h, w)       -- A.1, used by decode
 *   table 1  encode anchors   float32[N,4] (cy, cx, h', w')   -- A.2, used by the encoding
 *   table 2  anchor corners   float32[N,4] (ymin,xmin,ymax,xmax) second trip -- IoU, 4th encode output
 *   table 3  inside mask      uint8[N]                        -- ssd_common.py:112-115
 * sizes/ratios are ragged: layer l owns n_sizes[l] / n_ratios[l] consecutive doubles.
 * allowed_borders == NULL means "every anchor is inside" (SSD nets have no borders).
 */
int ronk_anchors_create(int kind, int img_h, int img_w, int num_layers,
                        const int* feat_shapes_host /*[L,2]*/,
                        const double* sizes_host, const int* n_sizes_host /*[L]*/,
                        const double* ratios_host, const int* n_ratios_host /*[L]*/,
                        const double* steps_host /*[L]*/, double offset,
                        const int* allowed_borders_host /*[L] or NULL*/,
                        ronk_anchors_t** out);
/* Arbitrary anchors, for callers of tf_ssd_bboxes_encode_layer / tf_ssd_bboxes_decode_layer
 * (nets/ssd_common.py:77-147,448-474) that bring their own flattened (y, x, h, w) float32[N,4]:
 * tables 0 and 1 are the given values, table 2 is one corner trip (ssd_common.py:105-108), the
 * inside mask uses per-anchor borders (int32[N] or NULL = all inside).  One layer of shape (N,1,1). */
int ronk_anchors_create_flat(int img_h, int img_w, int N, const float* yxhw_host,
                             const int* allowed_border_host, ronk_anchors_t** out);
void ronk_anchors_destroy(ronk_anchors_t* h);
int ronk_anchors_num(const ronk_anchors_t* h);                 /* N */
int ronk_anchors_num_layers(const ronk_anchors_t* h);          /* L */
int ronk_anchors_layer_info(const ronk_anchors_t* h, int layer, int* H, int* W, int* A, int* offset);
const void* ronk_anchors_table(const ronk_anchors_t* h, int which /*0..3*/); /* device pointer */
/* per-layer h[A], w[A] exactly as the reference stores them (float32), host copy */
int ronk_anchors_layer_hw(const ronk_anchors_t* h, int layer, float* h_host, float* w_host);

/* ------------------------------------------------------------ match + encode
 * Replaces RONNet.bboxes_encode -> tf_ssd_bboxes_encode -> tf_ssd_bboxes_encode_layer
 * -> iou_matrix + do_dual_max_match (nets/ron_vgg_320.py:173-186,
 * nets/ssd_common.py:337-414,77-147,27-75), batched over B images, plus the
 * objectness-prior label of ron_losses (nets/ron_vgg_320.py:686,708).
 * gt_boxes [B,Gmax,4], gt_labels int64 [B,Gmax], gt_counts int32 [B] (1 <= count <= Gmax).
 * out_labels int64 [B,N] in {-1 ignore, 0 negative, class}; out_loc [B,N,4] (cx,cy,w,h);
 * out_scores [B,N]; out_matched int32 [B,N] in {-2,-1,0..G-1} or NULL; out_objness int32 [B,N] or NULL.
 * The workspace must be zeroed once (ronk_encode_workspace_init); every call leaves it zeroed.
 */
size_t ronk_encode_workspace_bytes(int B, int Gmax);
int ronk_encode_workspace_init(void* ws, int B, int Gmax, void* stream);
int ronk_match_encode(const ronk_anchors_t* h,
                      const float* gt_boxes, const int64_t* gt_labels, const int32_t* gt_counts,
                      int B, int Gmax, float positive_threshold, float ignore_threshold,
                      const float* prior_scaling_host /*[4]*/, int match_flags,
                      int64_t* out_labels, float* out_loc, float* out_scores,
                      int32_t* out_matched, int32_t* out_objness,
                      void* ws, void* stream);

/* -------------------------------------------------------------------- decode
 * Replaces RONNet.bboxes_decode -> tf_ssd_bboxes_decode(_layer)
 * (nets/ron_vgg_320.py:188-195, nets/ssd_common.py:477-498,448-474).
 * loc [B,n,4] for anchors [first_anchor, first_anchor+n) -> boxes [B,n,4].
 */
int ronk_decode(const ronk_anchors_t* h, const float* loc, int B, int first_anchor, int n,
                const float* prior_scaling_host, float* out_boxes, void* stream);

/* ------------------------------------------- decode + filter + select + top-k
 * Replaces, fused and batched: bboxes_decode, the objectness gate of
 * eval_ron_network.py:227-229, tf_ssd_bboxes_select(_layer) (nets/ssd_common.py:552-589,
 * 504-549), tfe.bboxes_clip (tf_extended/bboxes.py:105-144), RONNet.bboxes_filter_min
 * (nets/ron_vgg_320.py:196-233) and tfe.bboxes_sort (tf_extended/bboxes.py:60-101).
 * Per-layer inputs exactly as the network emits them (no concatenation):
 *   loc_layers[l]  float32 [B, n_l, 4]   cls_layers[l] float32 [B, n_l, C]
 *   obj_layers[l]  float32 [B, n_l] or obj_layers == NULL (SSD / already gated)
 * (the *_layers arrays are HOST arrays of L device pointers, 16-byte aligned).
 * select_threshold >= 0 (None in the reference == 0).  clip_host == NULL: no clipping;
 * min_size < 0: no min-size filter (SSD order: select -> sort).
 * out_scores [B,C-1,K], out_boxes [B,C-1,K,4], out_idx int32 [B,C-1,K] (anchor index, -1 = padding) or NULL.
 */
size_t ronk_select_workspace_bytes(const ronk_anchors_t* h, int B, int C, int K);
int ronk_decode_select_topk(const ronk_anchors_t* h,
                            const float* const* loc_layers_host, const float* const* cls_layers_host,
                            const float* const* obj_layers_host,
                            int B, int C, float objectness_threshold, float select_threshold,
                            const float* clip_host /*[4] or NULL*/, float min_size,
                            const float* prior_scaling_host, int K, int select_flags,
                            float* out_scores, float* out_boxes, int32_t* out_idx,
                            void* ws, void* stream);

/* ------------------------------------------------------- generic sort (top-k)
 * Replaces tfe.bboxes_sort on arbitrary inputs (tf_extended/bboxes.py:60-101):
 * per row, tf.nn.top_k(scores, K) (ties: lower index first) + gather + pad_axis.
 * scores [S,N], boxes [S,N,4] -> out_scores [S,Kout], out_boxes [S,Kout,4], out_idx int32 or NULL,
 * Kout = K (rows are zero-padded when N < K, tensors.py:59-86).
 */
int ronk_sort_topk(const float* scores, const float* boxes, int S, int N, int K,
                   float* out_scores, float* out_boxes, int32_t* out_idx, void* stream);
/* The same per-row tf.nn.top_k(sorted=True) for rows of any length below 2^24 (ronk_sort_topk sorts in shared memory:
 * K <= 16 384; the reference's tf.nn.top_k(k = number of boxes), ron_eval.py:155 / :217 / :301, has no bound): a stable
 * LSD radix sort in global memory of one composite key per element (row, descending score, column).  S <= 256, K <= N;
 * boxes / out_boxes may both be NULL. */
size_t ronk_sort_rows_workspace_bytes(int S, int N);
int ronk_sort_rows(const float* scores, const float* boxes, int S, int N, int K, float* out_scores, float* out_boxes,
                   int32_t* out_idx, void* ws, size_t ws_bytes, void* stream);

/* ----------------------------------------------------------------------- clip
 * Replaces tfe.bboxes_clip (tf_extended/bboxes.py:105-144) on n boxes. */
int ronk_clip(const float* clip_host /*[4]*/, const float* boxes, long long n, float* out_boxes, void* stream);

/* ------------------------------------------------------------------------ NMS
 * Replaces tfe.bboxes_nms_batch -> bboxes_nms (tf_extended/bboxes.py:262-302,173-234).
 * scores [S,K], boxes [S,K,4]; rows are stably re-sorted by decreasing score unless
 * assume_sorted != 0.  out_scores [S,M], out_boxes [S,M,4] zero padded;
 * out_idx int32 [S,M] = position in the input row of each kept entry (-1 = padding) or NULL.
 */
size_t ronk_nms_workspace_bytes(int S, int K);
int ronk_nms_batch(const float* scores, const float* boxes, int S, int K,
                   float nms_threshold, int keep_top_k, int mode, int assume_sorted,
                   float* out_scores, float* out_boxes, int32_t* out_idx,
                   void* ws, void* stream);
/* Two-tier top-k for large K (the fused detect path): greedy NMS walks the sorted candidates and stops after keep_top_k
 * picks, so it rarely looks past the first few hundred.  Tier 1 = ronk_decode_select_topk with K1 = min(K, 1024) +
 * ronk_nms_batch_tiered(out_short): a row is marked when its K1 candidates ran out before keep_top_k boxes were kept.
 * Tier 2, for the marked rows only (the others return at once and keep their tier-1 outputs): ronk_select_topk_flagged
 * re-selects with the full K from the same workspace (an exact rebuild of the candidate list where the tier-1 pivot
 * cut it short) and ronk_nms_batch_tiered(only_flagged) redoes their NMS.  Results are those of one tier with K. */
int ronk_select_topk_flagged(const ronk_anchors_t* h,
                             const float* const* loc_layers_host, const float* const* cls_layers_host,
                             const float* const* obj_layers_host,
                             int B, int C, float objectness_threshold, float select_threshold,
                             const float* clip_host, float min_size, const float* prior_scaling_host, int K,
                             int select_flags, const int32_t* flags,
                             float* out_scores, float* out_boxes, int32_t* out_idx, void* ws, void* stream);
int ronk_nms_batch_tiered(const float* scores, const float* boxes, int S, int K, float nms_threshold,
                          int keep_top_k, int mode, const int32_t* only_flagged, int32_t* out_short,
                          float* out_scores, float* out_boxes, int32_t* out_idx, void* stream);

/* ---------------------------------------------------------------------- TP/FP
 * Replaces tfe.bboxes_matching_batch -> bboxes_matching -> bboxes_jaccard
 * (tf_extended/bboxes.py:407-450,316-404,527-554) for classes 1..C-1 at once.
 * det_scores [B,C-1,M], det_boxes [B,C-1,M,4]; glabels int64 [B,Gmax], gboxes [B,Gmax,4],
 * gdifficults int64 [B,Gmax] (zero padded like tf.train.batch(dynamic_pad=True)).
 * out_n_gt int64 [B,C-1], out_tp / out_fp uint8 [B,C-1,M].
 */
int ronk_tpfp_match(const float* det_scores, const float* det_boxes, int B, int C, int M,
                    const int64_t* glabels, const float* gboxes, const int64_t* gdifficults, int Gmax,
                    float matching_threshold,
                    int64_t* out_n_gt, uint8_t* out_tp, uint8_t* out_fp, void* stream);
/* Device-resident form of tfe.streaming_tp_fp_arrays (tf_extended/metrics.py:133-206): the detections of one batch
 * that pass the reference's filter (score > min_score, 1e-4 in the reference, and tp or fp, :167-175) are appended to
 * `records` (uint64 [capacity]: score bits << 32 | class index << 8 | fp << 1 | tp) in (class, image, rank) order
 * behind the earlier batches; n_gt_acc int64 [C-1] accumulates the ground-truth counts.  totals int32 [4] =
 * {count, count, overflow flag, -} must start zeroed; call_parity = number of earlier calls on these buffers (the two
 * count slots alternate between "before" and "after"); seg_counts int32 [C-1] (or NULL) receives this call's records per
 * class, which lets the host cut the class-major batches apart without sorting.  Nothing is read back: one copy at the
 * end of the evaluation. */
size_t ronk_tpfp_records_workspace_bytes(int B, int C, int M);
int ronk_tpfp_records_append(const float* det_scores, const uint8_t* tp, const uint8_t* fp, const int64_t* n_gt,
                             int B, int C, int M, float min_score, uint64_t* records, int capacity,
                             int32_t* totals, int call_parity, int64_t* n_gt_acc, int32_t* seg_counts, void* ws,
                             void* stream);

/* precision_recall + average_precision_voc07 / _voc12 (tf_extended/metrics.py:100-130, :237-258, :212-234; called
 * from eval_ron_network.py:262-324) on the device-resident records of ronk_tpfp_records_append, all classes in one
 * call: `records` uint64 [n] in concatenation order (rank-major after an all-gather, batch after batch); an entry whose
 * class index is 0xffffff is padding and ignored.  A stable radix sort by (class, descending score) reproduces
 * tf.nn.top_k's tie order (lower index first); cumulative sums are exact integers; precision / recall / the envelope /
 * the two sums are float64 with the reference's safe division and a fixed reduction order (VOC07 equals the host
 * NumPy value bit for bit, VOC12 to rounding).  n_gt int64 [C-1] (device).  thresholds: HOST pointer to the
 * n_thresholds (<= 16) recall levels of VOC07 (np.arange(0., 1.1, 0.1) in the reference).  Outputs (device):
 * out_ap07 / out_ap12 float64 [C-1]; optional out_offsets int32 [C] (first sorted record of every class, then the
 * number of real records), out_sorted uint64 [n], out_precision / out_recall float64 [n] (sorted order). */
size_t ronk_average_precision_workspace_bytes(long long n, int C);
int ronk_average_precision_records(const uint64_t* records, long long n, const int64_t* n_gt, int C,
                                   const double* thresholds, int n_thresholds, double* out_ap07, double* out_ap12,
                                   int32_t* out_offsets, uint64_t* out_sorted, double* out_precision,
                                   double* out_recall, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------- fine-grained functions
 * One kernel per small reference function, so the whole Python surface is on the GPU:
 * ronk_areas            areas                      nets/ssd_common.py:27-29     boxes [n,4] -> [n]
 * ronk_pairwise         intersection / iou_matrix  nets/ssd_common.py:30-47     a [G,4], b [N,4] -> [G,N]
 *                       (mode 0 = intersection, 1 = IoU with 0 where the union is 0)
 * ronk_overlap_ref      bboxes_jaccard / bboxes_intersection  tf_extended/bboxes.py:527-583
 *                       ref [1 or N,4] (device) vs boxes [N,4] -> [N] (mode 0 = jaccard, 1 = intersection;
 *                       2 / 3 = the NumPy flavours of nets/np_methods.py:187-227: plain IEEE quotients, union =
 *                       (vol_ref + vol_box) - inter, intersection relative to the reference box)
 * ronk_select_mask      tf_ssd_bboxes_select_layer nets/ssd_common.py:504-549   pred [B,n,C], boxes [B,n,4]
 *                       -> scores [B,C',n], boxes [B,C',n,4] zeroed where score <= thr (C' skips ignore_class)
 * ronk_dual_max_match   do_dual_max_match          nets/ssd_common.py:49-75     overlap [G,N] -> matched int64 [N], scores [N]
 */
int ronk_areas(const float* boxes, int n, float* out, void* stream);
int ronk_pairwise(const float* a, int G, const float* b, int N, int mode, float* out, void* stream);
int ronk_overlap_ref(const float* ref, int ref_n, const float* boxes, int N, int mode, float* out, void* stream);
int ronk_select_mask(const float* pred, const float* boxes, int B, int n, int C, float select_threshold,
                     int ignore_class, float* out_scores, float* out_boxes, void* stream);
/* mixed-class flavour:
 * ronk_select_all_classes  tf_ssd_bboxes_select_layer_all_classes  nets/ssd_common.py:592-628
 *                          pred [rows,C] -> classes int64 [rows], scores [rows]; use_threshold == 0 is the
 *                          reference's "select_threshold is None or 0" branch
 * ronk_gather_i64          the class gather of bboxes_sort_all_classes  tf_extended/bboxes.py:27-57 */
int ronk_select_all_classes(const float* pred, long long rows, int C, int use_threshold, float threshold,
                            int64_t* out_classes, float* out_scores, void* stream);
int ronk_gather_i64(const int64_t* src, const int32_t* idx, int S, int N, int K, int64_t* out, void* stream);
size_t ronk_dual_max_match_workspace_bytes(int G);
int ronk_dual_max_match(const float* overlap, int G, int N, float high_thres, float low_thres, int match_flags,
                        int64_t* out_matched, float* out_scores, void* ws, void* stream);

/* ------------------------------------------- ron_eval.py single-image variant
 * (the heavy steps reuse ronk_sort_topk + ronk_nms_batch, mode 'union')
 * ronk_flaten_predict     flaten_predict            ron_eval.py:111-144   per-layer pred [n_l,C], objness [n_l]
 *                         -> scores = objness * pred [N,C], labels = first arg-max int64 [N],
 *                            mask uint8 [N] = (label > 0) & (objness > threshold)   (batch size 1, like the reference)
 * ronk_filter_boxes_mask  filter_boxes              ron_eval.py:369-392   keep mask of boxes [n,4]
 * ronk_minsize_mask       RONNet.bboxes_filter_min  nets/ron_vgg_320.py:222-228   (w > minsize) & (h > minsize)
 * ronk_rowmax_mask        reduce_max + threshold    ron_eval.py:149-151   scores [n,C] -> max [n], mask [n]
 * ronk_compact_indices    tf.boolean_mask           order-preserving: indices of the set mask bytes + their count
 * ronk_gather_rows        tf.boolean_mask / gather  dst[i] = src[idx[i]] for rows of row_bytes (multiple of 4)
 * ronk_bboxes_resize      tfe.bboxes_resize         tf_extended/bboxes.py:147-171   (box - v) / s, bbox_ref on the host
 */
int ronk_flaten_predict(const float* const* pred_layers_host, const float* const* obj_layers_host,
                        const int* layer_sizes_host, int num_layers, int C, float objectness_threshold,
                        float* out_scores, int64_t* out_labels, uint8_t* out_mask, void* stream);
int ronk_filter_boxes_mask(const float* boxes, int n, float min_size, uint8_t* out_mask, void* stream);
int ronk_minsize_mask(const float* boxes, int n, float min_size, uint8_t* out_mask, void* stream);
/* RONNet.bboxes_filter_min (nets/ron_vgg_320.py:196-233) for S rows at once (all classes x images of the dict form):
 * ronk_filter_min_count gives, per row, how many boxes have width > min_size and height > min_size; with the padded
 * width = max(max count, top_k) known, ronk_filter_min_write writes the survivors of every row in order, zero-padded
 * (tf.boolean_mask + pad_axis).  scores [S,N], boxes [S,N,4] -> out_scores [S,width], out_boxes [S,width,4]. */
int ronk_filter_min_count(const float* boxes, int S, int N, float min_size, int32_t* out_counts, void* stream);
int ronk_filter_min_write(const float* scores, const float* boxes, int S, int N, float min_size, int width,
                          float* out_scores, float* out_boxes, void* stream);
int ronk_rowmax_mask(const float* scores, int n, int C, float threshold, float* out_max, uint8_t* out_mask,
                     void* stream);
size_t ronk_compact_workspace_bytes(int n);
int ronk_compact_indices(const uint8_t* mask, int n, int32_t* out_idx, int32_t* out_count, void* ws, void* stream);
int ronk_gather_rows(const void* src, int row_bytes, const int32_t* idx, int m, void* dst, void* stream);
int ronk_bboxes_resize(const float* bbox_ref_host /*[4]*/, const float* boxes, long long n, float* out_boxes,
                       void* stream);

/* per-class NMS variants of ron_eval.py (the heavy steps are ronk_sort_topk + ronk_nms_batch):
 * ronk_class_columns   tf_bboxes_nms_by_class  ron_eval.py:212-232   scores [n,C], boxes [n,4] -> one segment per class:
 *                      col_scores [C,n], col_boxes [C,n,4]; entries that do not start alive (score <= threshold, :228)
 *                      are zeroed, so they sort last and never suppress or get suppressed
 * ronk_keep_by_class   ron_eval.py:263-264,276-288   kept_pos / kept_scores [C,M] (positions in each class's sorted
 *                      list, -1 = padding), sorted_idx [C,n] -> max / first arg-max over classes of scores * keep_mask,
 *                      out_mask = max > 0; keep_ws is n*C bytes of scratch
 * ronk_group_by_label  tf_bboxes_nms_by_class_v1  ron_eval.py:338-345   sorted labels / scores / boxes -> segment c-1 holds,
 *                      in order, the entries with label c (c = 1..num_classes-1), zero padded to n; seg_pos = their
 *                      positions in the sorted list (-1 = padding)
 * ronk_mark_positions  ron_eval.py:344   OR of the per-class keep masks: out_mask[seg_pos[s, kept[s,m]]] = 1 */
int ronk_class_columns(const float* scores, const float* boxes, int n, int C, float threshold,
                       float* col_scores, float* col_boxes, void* stream);
int ronk_keep_by_class(const float* scores, int n, int C, const int32_t* kept_pos, const float* kept_scores, int M,
                       const int32_t* sorted_idx, float threshold, uint8_t* keep_ws, float* out_max,
                       int64_t* out_labels, uint8_t* out_mask, void* stream);
int ronk_group_by_label(const int64_t* labels, const float* scores, const float* boxes, int n, int num_classes,
                        float* seg_scores, float* seg_boxes, int32_t* seg_pos, void* stream);
int ronk_mark_positions(const int32_t* kept, const int32_t* seg_pos, int S, int M, int n, uint8_t* out_mask,
                        void* stream);

/* ------------------------------------------- RON loss example masks + smooth-L1
 * (the consumer of the encode outputs)
 * ronk_loss_masks         ron_losses  nets/ron_vgg_320.py:686-740   flat gclasses int64 [n], objectness scores [n] and
 *                         the two tf.random_uniform draws [n] (:705,:738; inputs, so results are reproducible) ->
 *                         final_neg_mask_objness u8 [n] (:707), objness_pred_label i32 [n] (:710),
 *                         cls_positive_mask u8 [n] (:726), final_cls_neg_mask_objness u8 [n] (:740),
 *                         counts f32 [4] or NULL = n_positives, n_negtives, n_cls_positives, n_cls_negtives.
 *                         n < 2^24 (the reference counts with float32 sums).  With localisations / glocalisations
 *                         [n,4] (or both NULL) the same launch also produces out_loss[0] = the localisation term below.
 *                         ws: ronk_loss_workspace_bytes(), zeroed once with ronk_loss_workspace_init; every call
 *                         leaves it zeroed again.  One cooperative launch
 * ronk_smooth_l1          modified_smooth_l1  nets/custom_layers.py:31-50   element-wise, scalar weights
 * ronk_localization_loss  nets/ron_vgg_320.py:760-764   beta * mean over cls_positive of the row sums of
 *                         modified_smooth_l1(sigma); 0 without positives.  Accumulates in double */
size_t ronk_loss_workspace_bytes(void);
int ronk_loss_workspace_init(void* ws, void* stream);
int ronk_loss_masks(const int64_t* gclasses, const float* objness_pred, const float* rand_objness,
                    const float* rand_cls, long long n, float objness_threshold, float negative_ratio,
                    uint8_t* out_final_objness, int32_t* out_objness_label, uint8_t* out_cls_positive,
                    uint8_t* out_final_cls, float* out_counts, const float* localisations,
                    const float* glocalisations, double sigma, float beta, float* out_loss, void* ws,
                    void* stream);
int ronk_smooth_l1(const float* pred, const float* target, long long count, float inside_weight,
                   float outside_weight, double sigma, float* out, void* stream);
int ronk_localization_loss(const float* localisations, const float* glocalisations, const uint8_t* cls_positive,
                           long long n, double sigma, float beta, float* out_loss, void* ws, void* stream);
/* Gradients of the two differentiable pieces (the reference differentiates the localisation term w.r.t. the network's
 * `localisations`; `glocalisations` is wrapped in tf.stop_gradient, nets/ron_vgg_320.py:760): d SmoothL1 / d pred
 * times grad_out element-wise, and d ronk_localization_loss / d localisations times the scalar *grad_out (device). */
int ronk_smooth_l1_backward(const float* pred, const float* target, const float* grad_out, long long count,
                            float inside_weight, float outside_weight, double sigma, float* grad_pred, void* stream);
int ronk_localization_loss_backward(const float* localisations, const float* glocalisations, const uint8_t* cls_positive,
                                    long long n, double sigma, float beta, const float* grad_out,
                                    float* grad_localisations, void* ws, void* stream);

/* ------------------------------------------- nets/np_methods.py (the notebooks' NumPy post-process)
 * (decode / sort / resize reuse ronk_decode, ronk_sort_topk, ronk_bboxes_resize)
 * ronk_np_select_mask    ssd_bboxes_select_layer  nets/np_methods.py:86-97   pred [n,C] -> keep mask over the
 *                        n*(C-1) (anchor, class >= 1) pairs in row-major order (pred > threshold), or over the n
 *                        anchors (argmax > 0) when use_threshold == 0 (select_threshold None / 0)
 * ronk_np_select_gather  same lines: classes int64 [m], scores [m], boxes [m,4] of the kept pairs idx[m]
 * ronk_np_clip           bboxes_clip  nets/np_methods.py:147-158   four max / min, bbox_ref on the host
 * ronk_np_nms            bboxes_nms   nets/np_methods.py:229-242 (bboxes_jaccard :181-201)   class-aware greedy
 *                        NMS on score-sorted boxes -> keep flags uint8 [n]; one launch, one CTA */
int ronk_np_select_mask(const float* pred, long long n, int C, int use_threshold, float threshold,
                        uint8_t* out_mask, void* stream);
int ronk_np_select_gather(const float* pred, const float* boxes, int C, int use_threshold, const int32_t* idx,
                          int m, int64_t* out_classes, float* out_scores, float* out_boxes, void* stream);
int ronk_np_clip(const float* bbox_ref_host /*[4]*/, const float* boxes, long long n, float* out_boxes, void* stream);
int ronk_np_nms(const int64_t* classes, const float* boxes, int n, float nms_threshold, uint8_t* out_keep,
                void* stream);

/* ------------------------------------------- datasets/voc_eval.py (offline PASCAL VOC evaluator)
 * 
 * ronk_voc_match   voc_eval  datasets/voc_eval.py:249-281   one class: detections float64 [nd,4] (x1,y1,x2,y2 as parsed
 *                  from the result file) sorted by decreasing confidence and grouped by image (det_offsets int32
 *                  [n_images+1]), ground truth float64 [ng,4] + difficult flags grouped the same way (gt_offsets),
 *                  max_gt = the largest group -> tp / fp uint8 [nd].  All arithmetic in float64 */
int ronk_voc_match(const double* det_boxes, const int32_t* det_offsets, const double* gt_boxes,
                   const int32_t* gt_offsets, const uint8_t* gt_difficult, int n_images, int max_gt,
                   double ovthresh, uint8_t* out_tp, uint8_t* out_fp, void* stream);

/* ------------------------------------------- host-buffer path of match + encode (sparse device -> host transfer)
 * 16 of the 28 target bytes per anchor are the localisation row, non-zero for ~1 % of the anchors:
 * ronk_sparse_rows_pack packs the non-zero rows of a device tensor rows [T,4] (T = B*N) into a fixed-capacity packet
 * {header int32[4] = (n, 0, 0, 0); row indices int32[cap]; pad to 16 B; rows float4[cap]} of
 * ronk_sparse_rows_packet_bytes(cap) bytes.  After the packet has been copied to the host, ronk_host_rows_apply
 * (plain host code) zeroes the rows the PREVIOUS packet of the same host array wrote (or pass NULL) and writes the new
 * ones; the array must be zero elsewhere.  Returns 1, touching nothing, when the packet overflowed (n > cap): copy
 * the dense tensor instead and zero the array before the next sparse step.  Labels and scores are copied dense. */
size_t ronk_sparse_rows_packet_bytes(int cap);
/* Labels of the same path: 0 (background) for ~95 % of the anchors.  ronk_sparse_labels_pack packs the non-zero
 * entries of labels int64 [T] as (index, int32 value) pairs into a packet of ronk_sparse_labels_packet_bytes(cap)
 * bytes.  ronk_host_targets_apply (plain host code, `threads` host threads from a persistent pool) applies a
 * localisation packet and a label packet to their host arrays in one go: it zeroes what the PREVIOUS packets wrote
 * (NULL: nothing), then writes the new entries.  It returns a bit mask and leaves the array in question untouched:
 * 1 = the localisation packet overflowed, 2 = the label packet overflowed or held a label outside int32.  Scores are
 * copied dense (a DMA write of 4 B per anchor is cheaper than any host-side scatter). */
size_t ronk_sparse_labels_packet_bytes(int cap);
int ronk_sparse_labels_pack(const int64_t* labels, long long T, int cap, void* packet_dev, void* stream);
int ronk_host_targets_apply(const void* loc_packet_host, const void* prev_loc_packet_host, int loc_cap, float* rows_host,
                            const void* lab_packet_host, const void* prev_lab_packet_host, int lab_cap,
                            int64_t* labels_host, int threads);
int ronk_sparse_rows_pack(const float* rows, long long T, int cap, void* packet_dev, void* stream);
int ronk_host_rows_apply(const void* packet_host, const void* prev_packet_host, int cap, float* rows_host);

/* number of kernel launches issued by this library in this process since load
 * (bench.py reports it as gpu_launches) */
long long ronk_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* RONK_H_ */
