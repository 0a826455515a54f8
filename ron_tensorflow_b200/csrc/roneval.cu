// Single-image post-process variant of the reference's ron_eval.py:
// flaten_predict (:111-144), filter_boxes (:369-392), the class-agnostic tf_bboxes_nms (:146-210)
// and tfe.bboxes_resize (tf_extended/bboxes.py:147-171).  The heavy pieces (sort, greedy NMS) are the
// kernels of postprocess.cu / nms.cu; this file adds the element-wise front ends and the
// order-preserving compaction (tf.boolean_mask) they all need:
//   flaten_kernel        scores = objness * predictions (one rounding), label = first arg-max,
//                        mask = label > 0 and objness > threshold, straight from the per-layer tensors;
//   box_filter_kernel    filter_boxes' keep mask;
//   rowmax_mask_kernel   reduce_max over classes + (max > select_threshold);
//   compact_count/compact_write   boolean_mask as a two-kernel stream compaction (per-tile counts,
//                        then every tile sums the counts before it and writes the kept indices);
//   gather_rows_kernel   rows of any width by index;
//   resize_kernel        (box - v) / s.
#include "common.cuh"

namespace ronk {

constexpr int kCompTile = 1024;

struct FlatenParams {
    const float* pred[kMaxLayers];
    const float* obj[kMaxLayers];
    int offs[kMaxLayers + 1];
    int L, C;
    float obj_thr;
    float* scores;      // [N, C]
    long long* labels;  // [N]
    uint8_t* mask;      // [N]
};

__global__ void __launch_bounds__(256)
flaten_kernel(const __grid_constant__ FlatenParams p) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int N = p.offs[p.L];
    if (n >= N) return;
    int l = 0;
    while (l + 1 < p.L && n >= p.offs[l + 1]) ++l;
    const int i = n - p.offs[l];
    const float o = p.obj[l][i];
    const float* row = p.pred[l] + (size_t)i * p.C;
    float* out = p.scores + (size_t)n * p.C;
    float best = 0.f;
    int label = 0;
    for (int c = 0; c < p.C; ++c) {
        const float v = o * row[c];                       // ron_eval.py:131 (expand_dims(objness) * pred)
        out[c] = v;
        if (c == 0 || v > best) { best = v; label = c; }  // tf.argmax: first occurrence (:134)
    }
    p.labels[n] = label;
    p.mask[n] = (label > 0 && o > p.obj_thr) ? 1 : 0;     // :141,143
}

__global__ void __launch_bounds__(256)
box_filter_kernel(const float4* __restrict__ boxes, int n, float min_size, uint8_t* __restrict__ mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 b = boxes[i];
    const float ws = b.w - b.y, hs = b.z - b.x;           // ron_eval.py:381-382
    const float x_ctr = b.y + ws / 2.f, y_ctr = b.x + hs / 2.f;
    mask[i] = (ws > min_size && hs > min_size && x_ctr > 0.f && y_ctr > 0.f && x_ctr < 1.f && y_ctr < 1.f) ? 1 : 0;
}

// RONNet.bboxes_filter_min (nets/ron_vgg_320.py:222-228): width > minsize and height > minsize
__global__ void __launch_bounds__(256)
minsize_mask_kernel(const float4* __restrict__ boxes, int n, float min_size, uint8_t* __restrict__ mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 b = boxes[i];
    mask[i] = ((b.w - b.y) > min_size && (b.z - b.x) > min_size) ? 1 : 0;
}

// RONNet.bboxes_filter_min for many rows at once (every (class, image) row of the dict form): a CTA per row counts the
// boxes that pass, then (second launch, once the padded width is known) writes them in order, zero-padded.
__global__ void __launch_bounds__(256)
filter_min_count_kernel(const float4* __restrict__ boxes, int N, float min_size, int* __restrict__ counts) {
    __shared__ int s_w[8];
    const float4* row = boxes + (size_t)blockIdx.x * N;
    int c = 0;
    for (int i = threadIdx.x; i < N; i += 256) {
        const float4 b = row[i];
        c += ((b.w - b.y) > min_size && (b.z - b.x) > min_size) ? 1 : 0;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += s_w[w];
        counts[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256)
filter_min_write_kernel(const float* __restrict__ scores, const float4* __restrict__ boxes, int N, float min_size, int width,
                        float* __restrict__ out_scores, float4* __restrict__ out_boxes) {
    __shared__ int s_w[8];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* srow = scores + (size_t)blockIdx.x * N;
    const float4* brow = boxes + (size_t)blockIdx.x * N;
    float* os = out_scores + (size_t)blockIdx.x * width;
    float4* ob = out_boxes + (size_t)blockIdx.x * width;
    int pos = 0;
    for (int i0 = 0; i0 < N; i0 += 256) {
        const int i = i0 + threadIdx.x;
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        bool keep = false;
        if (i < N) {
            b = brow[i];
            keep = (b.w - b.y) > min_size && (b.z - b.x) > min_size;
        }
        const unsigned m = __ballot_sync(full, keep);
        __syncthreads();
        if (lane == 0) s_w[warp] = __popc(m);
        __syncthreads();
        int before = 0, all = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            before += (w < warp) ? s_w[w] : 0;
            all += s_w[w];
        }
        if (keep) {
            const int o = pos + before + __popc(m & ((1u << lane) - 1u));
            os[o] = srow[i];
            ob[o] = b;
        }
        pos += all;
    }
    for (int o = pos + threadIdx.x; o < width; o += 256) {      // pad_axis (tensors.py:59-86)
        os[o] = 0.f;
        ob[o] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

__global__ void __launch_bounds__(256)
rowmax_mask_kernel(const float* __restrict__ scores, int n, int C, float thr, float* __restrict__ out,
                   uint8_t* __restrict__ mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* row = scores + (size_t)i * C;
    float m = row[0];
    for (int c = 1; c < C; ++c) m = fmaxf(m, row[c]);
    out[i] = m;
    mask[i] = m > thr ? 1 : 0;
}

__global__ void __launch_bounds__(256)
compact_count_kernel(const uint8_t* __restrict__ mask, int n, int* __restrict__ tile_counts) {
    __shared__ int s_w[8];
    const int base = blockIdx.x * kCompTile;
    int c = 0;
    for (int k = threadIdx.x; k < kCompTile; k += 256) {
        const int i = base + k;
        c += (i < n && mask[i]) ? 1 : 0;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += s_w[w];
        tile_counts[blockIdx.x] = t;
    }
}

// every tile sums the counts of the tiles before it (a few hundred at most), then writes the indices of
// its kept elements in order; the last tile also publishes the total
__global__ void __launch_bounds__(256)
compact_write_kernel(const uint8_t* __restrict__ mask, int n, const int* __restrict__ tile_counts, int tiles,
                     int* __restrict__ idx, int* __restrict__ total) {
    __shared__ int s_w[8];
    __shared__ int s_base;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int c = 0;
    for (int t = threadIdx.x; t < (int)blockIdx.x; t += 256) c += tile_counts[t];
    c = __reduce_add_sync(full, c);
    if (lane == 0) s_w[warp] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += s_w[w];
        s_base = t;
        if ((int)blockIdx.x == tiles - 1) *total = t + tile_counts[blockIdx.x];
    }
    __syncthreads();
    int pos = s_base;
    const int base = blockIdx.x * kCompTile;
    for (int k0 = 0; k0 < kCompTile; k0 += 256) {
        const int i = base + k0 + threadIdx.x;
        const bool keep = i < n && mask[i];
        const unsigned m = __ballot_sync(full, keep);
        __syncthreads();
        if (lane == 0) s_w[warp] = __popc(m);
        __syncthreads();
        int before = 0, all = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            before += (w < warp) ? s_w[w] : 0;
            all += s_w[w];
        }
        if (keep) idx[pos + before + __popc(m & ((1u << lane) - 1u))] = i;
        pos += all;
    }
}

__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ src, int row_floats, const int* __restrict__ idx, long long total,
                   float* __restrict__ dst) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const long long r = e / row_floats;
    const int c = (int)(e - r * row_floats);
    dst[e] = src[(size_t)idx[r] * row_floats + c];
}

__global__ void __launch_bounds__(256)
resize_kernel(const float4* __restrict__ in, long long n, float4 v, float4 s, float4* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 b = in[i];
    b.x = (b.x - v.x) / s.x;      // tf_extended/bboxes.py:163-170: translate, then one true division
    b.y = (b.y - v.y) / s.y;
    b.z = (b.z - v.z) / s.z;
    b.w = (b.w - v.w) / s.w;
    out[i] = b;
}


// ---- per-class NMS variants of ron_eval.py (tf_bboxes_nms_by_class :212-291, _v1 :293-366)

// by_class front end: column c of scores [n,C] becomes segment c; entries that do not start alive
// (score <= select_threshold, :228) become zero score / zero box, which sort after every live entry
// and never take part in a suppression (zero volume -> safe_divide gives 0).
__global__ void __launch_bounds__(256)
class_columns_kernel(const float* __restrict__ scores, const float4* __restrict__ boxes, int n, int C, float thr,
                     float* __restrict__ col_scores, float4* __restrict__ col_boxes) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)n * C) return;
    const int c = (int)(e / n), i = (int)(e - (long long)c * n);
    const float v = scores[(size_t)i * C + c];
    const bool live = v > thr;
    col_scores[e] = live ? v : 0.f;
    col_boxes[e] = live ? boxes[i] : make_float4(0.f, 0.f, 0.f, 0.f);
}

// keep_mask[orig, c] = 1 for every entry the NMS of class c kept (:263-264); kept_pos are positions
// in the sorted list of the class (-1 padding), sorted_idx maps them back to the rows.
__global__ void __launch_bounds__(256)
mark_kept_kernel(const int* __restrict__ kept_pos, const float* __restrict__ kept_scores,
                 const int* __restrict__ sorted_idx, int C, int M, int n, float thr, uint8_t* __restrict__ keep) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= C * M) return;
    const int c = e / M, pos = kept_pos[e];
    if (pos < 0 || !(kept_scores[e] > thr)) return;
    keep[(size_t)sorted_idx[(size_t)c * n + pos] * C + c] = 1;
}

// :282-288: keep_scores = scores * mask, its max / first arg-max over classes, keep = max > 0
__global__ void __launch_bounds__(256)
keep_reduce_kernel(const float* __restrict__ scores, const uint8_t* __restrict__ keep, int n, int C,
                   float* __restrict__ out_max, long long* __restrict__ out_label, uint8_t* __restrict__ out_mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float best = 0.f;
    int label = 0;
    for (int c = 0; c < C; ++c) {
        const float v = scores[(size_t)i * C + c] * (keep[(size_t)i * C + c] ? 1.f : 0.f);
        if (c == 0 || v > best) { best = v; label = c; }
    }
    out_max[i] = best;
    out_label[i] = label;
    out_mask[i] = best > 0.f ? 1 : 0;
}

// _v1 front end: one CTA per class c = blockIdx.x + 1 gathers, in order, the entries of the sorted
// list whose label is c (:340) into segment blockIdx.x; the rest of the segment is zero padding.
__global__ void __launch_bounds__(256)
group_by_label_kernel(const long long* __restrict__ labels, const float* __restrict__ scores,
                      const float4* __restrict__ boxes, int n, float* __restrict__ seg_scores,
                      float4* __restrict__ seg_boxes, int* __restrict__ seg_pos) {
    __shared__ int s_w[8];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long cls = (long long)blockIdx.x + 1;
    const size_t base = (size_t)blockIdx.x * n;
    int pos = 0;
    for (int k0 = 0; k0 < n; k0 += 256) {
        const int i = k0 + threadIdx.x;
        const bool mine = i < n && labels[i] == cls;
        const unsigned m = __ballot_sync(full, mine);
        __syncthreads();
        if (lane == 0) s_w[warp] = __popc(m);
        __syncthreads();
        int before = 0, all = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            before += (w < warp) ? s_w[w] : 0;
            all += s_w[w];
        }
        if (mine) {
            const int o = pos + before + __popc(m & ((1u << lane) - 1u));
            seg_scores[base + o] = scores[i];
            seg_boxes[base + o] = boxes[i];
            seg_pos[base + o] = i;
        }
        pos += all;
    }
    for (int o = pos + threadIdx.x; o < n; o += 256) {
        seg_scores[base + o] = 0.f;
        seg_boxes[base + o] = make_float4(0.f, 0.f, 0.f, 0.f);
        seg_pos[base + o] = -1;
    }
}

// _v1: total_keep_mask |= keep_mask of the class (:344); kept are positions inside the segment
__global__ void __launch_bounds__(256)
mark_positions_kernel(const int* __restrict__ kept, const int* __restrict__ seg_pos, int S, int M, int n,
                      uint8_t* __restrict__ mask) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= S * M) return;
    const int k = kept[e];
    if (k < 0) return;
    const int i = seg_pos[(size_t)(e / M) * n + k];
    if (i >= 0) mask[i] = 1;
}

}  // namespace ronk

using namespace ronk;

extern "C" int ronk_flaten_predict(const float* const* pred_layers, const float* const* obj_layers,
                                   const int* layer_sizes, int num_layers, int C, float objectness_threshold,
                                   float* out_scores, int64_t* out_labels, uint8_t* out_mask, void* stream) {
    RONK_REQUIRE(pred_layers && obj_layers && layer_sizes && out_scores && out_labels && out_mask, RONK_EINVAL,
                 "ronk_flaten_predict: NULL argument");
    RONK_REQUIRE(num_layers >= 1 && num_layers <= kMaxLayers && C >= 1, RONK_EINVAL,
                 "ronk_flaten_predict: need 1..16 layers and C >= 1");
    FlatenParams p;
    int n = 0;
    for (int l = 0; l < num_layers; ++l) {
        RONK_REQUIRE(pred_layers[l] && obj_layers[l] && layer_sizes[l] >= 0, RONK_EINVAL,
                     "ronk_flaten_predict: bad layer");
        p.pred[l] = pred_layers[l];
        p.obj[l] = obj_layers[l];
        p.offs[l] = n;
        n += layer_sizes[l];
    }
    p.offs[num_layers] = n;
    p.L = num_layers;
    p.C = C;
    p.obj_thr = objectness_threshold;
    p.scores = out_scores;
    p.labels = (long long*)out_labels;
    p.mask = out_mask;
    if (n == 0) return RONK_OK;
    flaten_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p);
    RONK_LAUNCHED();
    return RONK_OK;
}

extern "C" int ronk_filter_boxes_mask(const float* boxes, int n, float min_size, uint8_t* out_mask, void* stream) {
    RONK_REQUIRE(n >= 0, RONK_EINVAL, "ronk_filter_boxes_mask: bad argument");
    if (n == 0) return RONK_OK;                     // empty inputs may come with NULL pointers
    RONK_REQUIRE(boxes && out_mask && ((uintptr_t)boxes % 16) == 0, RONK_EINVAL,
                 "ronk_filter_boxes_mask: bad argument");
    box_filter_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const float4*)boxes, n, min_size, out_mask);
    RONK_LAUNCHED();
    return RONK_OK;
}

extern "C" int ronk_minsize_mask(const float* boxes, int n, float min_size, uint8_t* out_mask, void* stream) {
    RONK_REQUIRE(n >= 0, RONK_EINVAL, "ronk_minsize_mask: bad argument");
    if (n == 0) return RONK_OK;
    RONK_REQUIRE(boxes && out_mask && ((uintptr_t)boxes % 16) == 0, RONK_EINVAL, "ronk_minsize_mask: bad argument");
    minsize_mask_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const float4*)boxes, n, min_size, out_mask);
    RONK_LAUNCHED();
    return RONK_OK;
}

extern "C" int ronk_filter_min_count(const float* boxes, int S, int N, float min_size, int32_t* out_counts, void* stream) {
    RONK_REQUIRE(boxes && out_counts && S >= 1 && N >= 1 && ((uintptr_t)boxes % 16) == 0, RONK_EINVAL,
                 "ronk_filter_min_count: bad argument");
    filter_min_count_kernel<<<S, 256, 0, (cudaStream_t)stream>>>((const float4*)boxes, N, min_size, out_counts);
    RONK_LAUNCHED();
    return RONK_OK;
}

extern "C" int ronk_filter_min_write(const float* scores, const float* boxes, int S, int N, float min_size, int width,
                                     float* out_scores, float* out_boxes, void* stream) {
    RONK_REQUIRE(scores && boxes && out_scores && out_boxes && S >= 1 && N >= 1 && width >= 1, RONK_EINVAL,
                 "ronk_filter_min_write: bad argument");
    RONK_REQUIRE(((uintptr_t)boxes % 16) == 0 && ((uintptr_t)out_boxes % 16) == 0, RONK_EINVAL,
                 "ronk_filter_min_write: box pointers must be 16-byte aligned");
    filter_min_write_kernel<<<S, 256, 0, (cudaStream_t)stream>>>(scores, (const float4*)boxes, N, min_size, width, out_scores,
                                                                (float4*)out_boxes);
    RONK_LAUNCHED();
    return RONK_OK;
}

extern "C" int ronk_rowmax_mask(const float* scores, int n, int C, float threshold, float* out_max, uint8_t* out_mask,
                                void* stream) {
    RONK_REQUIRE(n >= 0 && C >= 1, RONK_EINVAL, "ronk_rowmax_mask: bad argument");
    if (n == 0) return RONK_OK;
    RONK_REQUIRE(scores && out_max && out_mask, RONK_EINVAL, "ronk_rowmax_mask: bad argument");
    rowmax_mask_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(scores, n, C, threshold, out_max, out_mask);
    RONK_LAUNCHED();
    return RONK_OK;
}

extern "C" size_t ronk_compact_workspace_bytes(int n) {
    return n < 1 ? 4 : (size_t)((n + kCompTile - 1) / kCompTile) * 4;
}

extern "C" int ronk_compact_indices(const uint8_t* mask, int n, int32_t* out_idx, int32_t* out_count, void* ws,
                                    void* stream) {
    RONK_REQUIRE(out_count && n >= 0 && (n == 0 || (mask && out_idx && ws)), RONK_EINVAL,
                 "ronk_compact_indices: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        RONK_CUDA(cudaMemsetAsync(out_count, 0, 4, st));
        return RONK_OK;
    }
    const int tiles = (n + kCompTile - 1) / kCompTile;
    compact_count_kernel<<<tiles, 256, 0, st>>>(mask, n, (int*)ws);
    RONK_LAUNCHED();
    compact_write_kernel<<<tiles, 256, 0, st>>>(mask, n, (const int*)ws, tiles, out_idx, out_count);
    RONK_LAUNCHED();
    return RONK_OK;
}

extern "C" int ronk_gather_rows(const void* src, int row_bytes, const int32_t* idx, int m, void* dst, void* stream) {
    RONK_REQUIRE(m >= 0 && row_bytes >= 4 && row_bytes % 4 == 0 && (m == 0 || (src && idx && dst)), RONK_EINVAL,
                 "ronk_gather_rows: bad argument (rows are multiples of 4 bytes)");
    if (m == 0) return RONK_OK;
    const long long total = (long long)m * (row_bytes / 4);
    gather_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const float*)src, row_bytes / 4, idx, total, (float*)dst);
    RONK_LAUNCHED();
    return RONK_OK;
}

extern "C" int ronk_bboxes_resize(const float* bbox_ref, const float* boxes, long long n, float* out_boxes,
                                  void* stream) {
    RONK_REQUIRE(bbox_ref && n >= 0 && (n == 0 || (boxes && out_boxes)), RONK_EINVAL, "ronk_bboxes_resize: bad argument");
    if (n == 0) return RONK_OK;
    RONK_REQUIRE(((uintptr_t)boxes % 16) == 0 && ((uintptr_t)out_boxes % 16) == 0, RONK_EINVAL,
                 "ronk_bboxes_resize: pointers must be 16-byte aligned");
    const float4 v = make_float4(bbox_ref[0], bbox_ref[1], bbox_ref[0], bbox_ref[1]);
    const float sh = bbox_ref[2] - bbox_ref[0], sw = bbox_ref[3] - bbox_ref[1];
    resize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)boxes, n, v,
                                                                              make_float4(sh, sw, sh, sw),
                                                                              (float4*)out_boxes);
    RONK_LAUNCHED();
    return RONK_OK;
}

extern "C" int ronk_class_columns(const float* scores, const float* boxes, int n, int C, float threshold,
                                  float* col_scores, float* col_boxes, void* stream) {
    RONK_REQUIRE(n >= 0 && C >= 1 && (n == 0 || (scores && boxes && col_scores && col_boxes)), RONK_EINVAL,
                 "ronk_class_columns: bad argument");
    if (n == 0) return RONK_OK;
    RONK_REQUIRE(((uintptr_t)boxes % 16) == 0 && ((uintptr_t)col_boxes % 16) == 0, RONK_EINVAL,
                 "ronk_class_columns: box pointers must be 16-byte aligned");
    const long long total = (long long)n * C;
    class_columns_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        scores, (const float4*)boxes, n, C, threshold, col_scores, (float4*)col_boxes);
    RONK_LAUNCHED();
    return RONK_OK;
}

extern "C" int ronk_keep_by_class(const float* scores, int n, int C, const int32_t* kept_pos, const float* kept_scores,
                                  int M, const int32_t* sorted_idx, float threshold, uint8_t* keep_ws, float* out_max,
                                  int64_t* out_labels, uint8_t* out_mask, void* stream) {
    RONK_REQUIRE(n >= 1 && C >= 1 && M >= 1 && scores && kept_pos && kept_scores && sorted_idx && keep_ws && out_max &&
                     out_labels && out_mask,
                 RONK_EINVAL, "ronk_keep_by_class: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    RONK_CUDA(cudaMemsetAsync(keep_ws, 0, (size_t)n * C, st));
    mark_kept_kernel<<<(unsigned)((C * M + 255) / 256), 256, 0, st>>>(kept_pos, kept_scores, sorted_idx, C, M, n, threshold,
                                                                  keep_ws);
    RONK_LAUNCHED();
    keep_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(scores, keep_ws, n, C, out_max, (long long*)out_labels,
                                                                 out_mask);
    RONK_LAUNCHED();
    return RONK_OK;
}

extern "C" int ronk_group_by_label(const int64_t* labels, const float* scores, const float* boxes, int n,
                                   int num_classes, float* seg_scores, float* seg_boxes, int32_t* seg_pos, void* stream) {
    RONK_REQUIRE(n >= 1 && num_classes >= 2 && labels && scores && boxes && seg_scores && seg_boxes && seg_pos, RONK_EINVAL,
                 "ronk_group_by_label: bad argument");
    RONK_REQUIRE(((uintptr_t)boxes % 16) == 0 && ((uintptr_t)seg_boxes % 16) == 0, RONK_EINVAL,
                 "ronk_group_by_label: box pointers must be 16-byte aligned");
    group_by_label_kernel<<<(unsigned)(num_classes - 1), 256, 0, (cudaStream_t)stream>>>(
        (const long long*)labels, scores, (const float4*)boxes, n, seg_scores, (float4*)seg_boxes, seg_pos);
    RONK_LAUNCHED();
    return RONK_OK;
}

extern "C" int ronk_mark_positions(const int32_t* kept, const int32_t* seg_pos, int S, int M, int n, uint8_t* out_mask,
                                   void* stream) {
    RONK_REQUIRE(S >= 1 && M >= 1 && n >= 1 && kept && seg_pos && out_mask, RONK_EINVAL, "ronk_mark_positions: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    RONK_CUDA(cudaMemsetAsync(out_mask, 0, (size_t)n, st));
    mark_positions_kernel<<<(unsigned)((S * M + 255) / 256), 256, 0, st>>>(kept, seg_pos, S, M, n, out_mask);
    RONK_LAUNCHED();
    return RONK_OK;
}
