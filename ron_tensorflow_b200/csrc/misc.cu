// Fine-grained entry points that mirror the reference's small Python functions one to one
// (nets/ssd_common.py:27-75,504-549 and tf_extended/bboxes.py:527-583).  The fused kernels in
// match_encode.cu / postprocess.cu are the hot path; these exist so that every function of the
// reference's surface has a CUDA implementation with identical results.
#include "common.cuh"

namespace ronk {

// areas (ssd_common.py:27-29)
__global__ void __launch_bounds__(256) areas_kernel(const float4* __restrict__ b, int n, float* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (b[i].w - b[i].y) * (b[i].z - b[i].x);
}

// intersection / iou_matrix (ssd_common.py:30-47): a [G,4] x b [N,4] -> [G,N]
__global__ void __launch_bounds__(256)
pairwise_kernel(const float4* __restrict__ a, int G, const float4* __restrict__ b, int N, int mode,
                float* __restrict__ out) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    int g = blockIdx.y;
    if (n >= N) return;
    float4 x = a[g], y = b[n];
    float h = fmaxf(fminf(x.z, y.z) - fmaxf(x.x, y.x), 0.f);
    float w = fmaxf(fminf(x.w, y.w) - fmaxf(x.y, y.y), 0.f);
    float inter = h * w;
    float r = inter;
    if (mode == 1) {
        float uni = ((x.w - x.y) * (x.z - x.x) + (y.w - y.y) * (y.z - y.x)) - inter;
        r = div_overlap(inter, uni);
    }
    out[(size_t)g * N + n] = r;
}

// bboxes_jaccard / bboxes_intersection (tf_extended/bboxes.py:527-583): ref [1 or N,4] vs boxes [N,4]
__global__ void __launch_bounds__(256)
overlap_ref_kernel(const float4* __restrict__ ref, int ref_n, const float4* __restrict__ boxes, int N, int mode,
                   float* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float4 r = ref[ref_n == 1 ? 0 : i], b = boxes[i];
    float h = fmaxf(fminf(b.z, r.z) - fmaxf(b.x, r.x), 0.f);
    float w = fmaxf(fminf(b.w, r.w) - fmaxf(b.y, r.y), 0.f);
    float inter = h * w;
    float bvol = (b.z - b.x) * (b.w - b.y);
    float rvol = (r.z - r.x) * (r.w - r.y);
    if (mode >= 2) {
        // nets/np_methods.py:187-227: plain IEEE quotients, union = (vol_ref + vol_box) - inter, the
        // intersection score is relative to the REFERENCE box
        out[i] = __fdiv_rn(inter, mode == 2 ? ((rvol + bvol) - inter) : rvol);
        return;
    }
    float den = (mode == 0) ? ((-inter + bvol) + rvol) : bvol;
    out[i] = (den > 0.f) ? inter / den : 0.f;   // tf_extended/math.py:25-38
}

// tf_ssd_bboxes_select_layer (ssd_common.py:537-547): scores * mask, boxes * mask per class
__global__ void __launch_bounds__(256)
select_mask_kernel(const float* __restrict__ pred, const float4* __restrict__ boxes, int B, int n, int C, float thr,
                   int ignore_class, float* __restrict__ out_s, float4* __restrict__ out_b) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over B*n
    if (i >= (long long)B * n) return;
    int b = (int)(i / n), a = (int)(i % n);
    float4 bx = boxes[i];
    int k = 0;
    int CM = (ignore_class >= 0 && ignore_class < C) ? C - 1 : C;
    for (int c = 0; c < C; ++c) {
        if (c == ignore_class) continue;
        float s = pred[i * C + c];
        float m = (s > thr) ? 1.f : 0.f;
        size_t o = ((size_t)b * CM + k) * n + a;
        out_s[o] = s * m;
        out_b[o] = make_float4(bx.x * m, bx.y * m, bx.z * m, bx.w * m);
        ++k;
    }
}

// tf_ssd_bboxes_select_layer_all_classes (ssd_common.py:592-628): one class + score per anchor.
// use_thr == 0 (select_threshold None or 0): arg-max / max over ALL classes, score zeroed for class 0;
// else: arg-max / max over classes 1.., class and score zeroed unless score > threshold.
__global__ void __launch_bounds__(256)
select_all_kernel(const float* __restrict__ pred, long long rows, int C, int use_thr, float thr,
                  long long* __restrict__ out_cls, float* __restrict__ out_s) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const float* row = pred + i * C;
    const int c0 = use_thr ? 1 : 0;
    float best = row[c0];
    int arg = c0;
    for (int c = c0 + 1; c < C; ++c) {
        const float v = row[c];
        if (v > best) { best = v; arg = c; }      // tf.argmax: first occurrence
    }
    float m;
    if (use_thr) m = (best > thr) ? 1.f : 0.f;    // :621-623
    else m = (arg > 0) ? 1.f : 0.f;               // :616
    out_cls[i] = use_thr ? (long long)arg * (long long)m : (long long)arg;
    out_s[i] = best * m;
}

// out[s, k] = src[s, idx[s, k]] for int64 rows (bboxes_sort_all_classes, tf_extended/bboxes.py:44-54); idx < 0 -> 0
__global__ void __launch_bounds__(256)
gather_i64_kernel(const long long* __restrict__ src, const int* __restrict__ idx, int S, int N, int K,
                  long long* __restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)S * K) return;
    const long long srow = e / K;
    const int j = idx[e];
    out[e] = j >= 0 ? src[srow * N + j] : 0ll;
}

__device__ __forceinline__ unsigned ord_bits(float s) {
    unsigned u = __float_as_uint(s);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// do_dual_max_match (ssd_common.py:49-75) on an explicit overlap matrix.
__global__ void __launch_bounds__(256)
dmm_rows_kernel(const float* __restrict__ ov, int N, int* __restrict__ g2a) {
    __shared__ u64 s_key[8];
    const int g = blockIdx.x;
    u64 best = 0ull;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        u64 k = ((u64)ord_bits(ov[(size_t)g * N + n]) << 32) | (u64)(0xffffffffu - (unsigned)n);
        best = k > best ? k : best;
    }
    for (int o = 16; o > 0; o >>= 1) {
        u64 v = __shfl_xor_sync(0xffffffffu, best, o);
        best = v > best ? v : best;
    }
    if ((threadIdx.x & 31) == 0) s_key[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) best = s_key[w] > best ? s_key[w] : best;
        g2a[g] = (int)(0xffffffffu - (unsigned)(best & 0xffffffffull));
    }
}

__global__ void __launch_bounds__(256)
dmm_cols_kernel(const float* __restrict__ ov, int G, int N, float high, float low, int ignore_between,
                int gt_max_first, long long* __restrict__ matched, float* __restrict__ scores,
                int* __restrict__ claimed) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float mv = ov[n];
    int a2g = 0;
    for (int g = 1; g < G; ++g) {
        float v = ov[(size_t)g * N + n];
        if (v > mv) { mv = v; a2g = g; }
    }
    bool less = mv < low;
    bool between = (mv < high) && (mv >= low);
    bool neg = ignore_between ? less : between;
    bool ign = ignore_between ? between : less;
    int mi = ign ? -2 : (neg ? -1 : a2g);
    matched[n] = mi;
    scores[n] = mv;
    if (!gt_max_first && mi >= 0) atomicOr(claimed + mi, 1);
}

__global__ void __launch_bounds__(256)
dmm_force_kernel(const float* __restrict__ ov, int G, int N, const int* __restrict__ g2a,
                 const int* __restrict__ claimed, int gt_max_first, long long* __restrict__ matched,
                 float* __restrict__ scores) {
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        if (!gt_max_first && claimed[g]) continue;
        int n = g2a[g];
        bool first = true;
        for (int g2 = 0; g2 < g; ++g2)
            if (g2a[g2] == n && (gt_max_first || !claimed[g2])) { first = false; break; }
        if (!first) continue;
        matched[n] = g;
        scores[n] = ov[(size_t)g * N + n];
    }
}

__global__ void zero_i32b_kernel(int* p, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0;
}

}  // namespace ronk

using namespace ronk;

extern "C" int ronk_areas(const float* boxes, int n, float* out, void* stream) {
    RONK_REQUIRE(boxes && out && n >= 1 && ((uintptr_t)boxes % 16) == 0, RONK_EINVAL, "ronk_areas: bad argument");
    areas_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const float4*)boxes, n, out);
    RONK_LAUNCHED();
    return RONK_OK;
}

extern "C" int ronk_pairwise(const float* a, int G, const float* b, int N, int mode, float* out, void* stream) {
    RONK_REQUIRE(a && b && out && G >= 1 && N >= 1 && G <= 65535 && (mode == 0 || mode == 1), RONK_EINVAL,
                 "ronk_pairwise: bad argument");
    RONK_REQUIRE(((uintptr_t)a % 16) == 0 && ((uintptr_t)b % 16) == 0, RONK_EINVAL, "ronk_pairwise: 16-byte alignment");
    pairwise_kernel<<<dim3((N + 255) / 256, G), 256, 0, (cudaStream_t)stream>>>((const float4*)a, G, (const float4*)b,
                                                                            N, mode, out);
    RONK_LAUNCHED();
    return RONK_OK;
}

extern "C" int ronk_overlap_ref(const float* ref, int ref_n, const float* boxes, int N, int mode, float* out,
                                void* stream) {
    RONK_REQUIRE(ref && boxes && out && N >= 1 && (ref_n == 1 || ref_n == N) && mode >= 0 && mode <= 3, RONK_EINVAL,
                 "ronk_overlap_ref: bad argument");
    RONK_REQUIRE(((uintptr_t)ref % 16) == 0 && ((uintptr_t)boxes % 16) == 0, RONK_EINVAL,
                 "ronk_overlap_ref: 16-byte alignment");
    overlap_ref_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const float4*)ref, ref_n,
                                                                      (const float4*)boxes, N, mode, out);
    RONK_LAUNCHED();
    return RONK_OK;
}

extern "C" int ronk_select_mask(const float* pred, const float* boxes, int B, int n, int C, float thr,
                                int ignore_class, float* out_scores, float* out_boxes, void* stream) {
    RONK_REQUIRE(pred && boxes && out_scores && out_boxes && B >= 1 && n >= 1 && C >= 1, RONK_EINVAL,
                 "ronk_select_mask: bad argument");
    RONK_REQUIRE(((uintptr_t)boxes % 16) == 0 && ((uintptr_t)out_boxes % 16) == 0, RONK_EINVAL,
                 "ronk_select_mask: 16-byte alignment");
    long long tot = (long long)B * n;
    select_mask_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        pred, (const float4*)boxes, B, n, C, thr, ignore_class, out_scores, (float4*)out_boxes);
    RONK_LAUNCHED();
    return RONK_OK;
}

extern "C" int ronk_select_all_classes(const float* pred, long long rows, int C, int use_threshold, float threshold,
                                       int64_t* out_classes, float* out_scores, void* stream) {
    RONK_REQUIRE(rows >= 0 && C >= 1 && (!use_threshold || C >= 2), RONK_EINVAL, "ronk_select_all_classes: bad sizes");
    if (rows == 0) return RONK_OK;
    RONK_REQUIRE(pred && out_classes && out_scores, RONK_EINVAL, "ronk_select_all_classes: NULL argument");
    select_all_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        pred, rows, C, use_threshold ? 1 : 0, threshold, (long long*)out_classes, out_scores);
    RONK_LAUNCHED();
    return RONK_OK;
}

extern "C" int ronk_gather_i64(const int64_t* src, const int32_t* idx, int S, int N, int K, int64_t* out, void* stream) {
    RONK_REQUIRE(S >= 0 && N >= 0 && K >= 0, RONK_EINVAL, "ronk_gather_i64: bad sizes");
    if ((long long)S * K == 0) return RONK_OK;
    RONK_REQUIRE(src && idx && out, RONK_EINVAL, "ronk_gather_i64: NULL argument");
    gather_i64_kernel<<<(unsigned)(((long long)S * K + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const long long*)src, idx, S, N, K, (long long*)out);
    RONK_LAUNCHED();
    return RONK_OK;
}

extern "C" size_t ronk_dual_max_match_workspace_bytes(int G) { return G < 1 ? 0 : (size_t)G * 8; }

extern "C" int ronk_dual_max_match(const float* overlap, int G, int N, float high, float low, int match_flags,
                                   int64_t* out_matched, float* out_scores, void* ws, void* stream) {
    RONK_REQUIRE(overlap && out_matched && out_scores && ws && G >= 1 && N >= 1, RONK_EINVAL,
                 "ronk_dual_max_match: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    int* g2a = (int*)ws;
    int* claimed = g2a + G;
    int ib = (match_flags & RONK_MATCH_NO_IGNORE_BETWEEN) ? 0 : 1;
    int gf = (match_flags & RONK_MATCH_NO_GT_MAX_FIRST) ? 0 : 1;
    zero_i32b_kernel<<<(G + 255) / 256, 256, 0, st>>>(claimed, G);
    RONK_LAUNCHED();
    dmm_rows_kernel<<<G, 256, 0, st>>>(overlap, N, g2a);
    RONK_LAUNCHED();
    dmm_cols_kernel<<<(N + 255) / 256, 256, 0, st>>>(overlap, G, N, high, low, ib, gf, (long long*)out_matched,
                                                      out_scores, claimed);
    RONK_LAUNCHED();
    dmm_force_kernel<<<1, 256, 0, st>>>(overlap, G, N, g2a, claimed, gf, (long long*)out_matched, out_scores);
    RONK_LAUNCHED();
    return RONK_OK;
}
