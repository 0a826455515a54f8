// The NumPy twin of the post-process that the reference's notebooks / demo use:
// nets/np_methods.py  ssd_bboxes_select_layer :58-100, bboxes_clip :147-158, bboxes_nms :229-242
// (with bboxes_jaccard :181-201).  Decode, sort and resize reuse ronk_decode / ronk_sort_topk /
// ronk_bboxes_resize; this file adds what differs from the TF graph flavour:
//   np_select_mask_kernel    the keep mask of np.where(pred[:, 1:] > thr) over (anchor, class) pairs in
//                            row-major order (an anchor yields one detection per class above the threshold),
//                            or argmax > 0 per anchor when the threshold is None / 0;
//   np_select_gather_kernel  (class, score, box) of every kept pair;
//   np_clip_kernel           four max / min against the reference box, nothing else;
//   np_nms_kernel            class-aware greedy NMS on the score-sorted list: box i, if still kept, drops every
//                            later box of ITS class whose jaccard with it is not < threshold.  The quotient is a
//                            plain IEEE division (0 / 0 = NaN is "not <", so identical empty boxes of one class
//                            are dropped), union = (vol_i + vol_j) - inter in that order.
#include "common.cuh"

namespace ronk {

__global__ void __launch_bounds__(256)
np_select_mask_kernel(const float* __restrict__ pred, long long n, int C, int use_threshold, float thr,
                      uint8_t* __restrict__ mask) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (use_threshold) {
        if (e >= n * (C - 1)) return;
        const long long row = e / (C - 1);
        const int c = (int)(e - row * (C - 1)) + 1;
        mask[e] = pred[row * C + c] > thr ? 1 : 0;                 // np_methods.py:94-95
    } else {
        if (e >= n) return;
        const float* r = pred + e * C;
        float best = r[0];
        int cls = 0;
        for (int c = 1; c < C; ++c)
            if (r[c] > best) { best = r[c]; cls = c; }               // np.argmax: first occurrence
        mask[e] = cls > 0 ? 1 : 0;                                   // :88-90
    }
}

__global__ void __launch_bounds__(256)
np_select_gather_kernel(const float* __restrict__ pred, const float4* __restrict__ boxes, int C, int use_threshold,
                        const int* __restrict__ idx, int m, long long* __restrict__ classes,
                        float* __restrict__ scores, float4* __restrict__ out_boxes) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    const int e = idx[k];
    int row, cls;
    float s;
    if (use_threshold) {
        row = e / (C - 1);
        cls = e - row * (C - 1) + 1;                                 // idxes[-1] + 1   (:96)
        s = pred[(size_t)row * C + cls];
    } else {
        row = e;
        const float* r = pred + (size_t)row * C;
        s = r[0];
        cls = 0;
        for (int c = 1; c < C; ++c)
            if (r[c] > s) { s = r[c]; cls = c; }
    }
    classes[k] = cls;
    scores[k] = s;
    out_boxes[k] = boxes[row];
}

__global__ void __launch_bounds__(256)
np_clip_kernel(const float4* __restrict__ in, long long n, float4 ref, float4* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 b = in[i];
    b.x = fmaxf(b.x, ref.x);                                         // np_methods.py:153-156
    b.y = fmaxf(b.y, ref.y);
    b.z = fminf(b.z, ref.z);
    b.w = fminf(b.w, ref.w);
    out[i] = b;
}

constexpr int kNpNmsThreads = 256;

// one CTA: keep flags in shared memory, the reference's outer loop over i runs in order
__global__ void __launch_bounds__(kNpNmsThreads)
np_nms_kernel(const long long* __restrict__ classes, const float4* __restrict__ boxes, int n, float thr,
              uint8_t* __restrict__ keep_out) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint8_t* s_keep = smem;
    for (int j = threadIdx.x; j < n; j += kNpNmsThreads) s_keep[j] = 1;
    __syncthreads();
    for (int i = 0; i + 1 < n; ++i) {                                // range(scores.size - 1)   (:233)
        if (s_keep[i]) {                                             // uniform: written before the last barrier
            const float4 bi = boxes[i];
            const long long ci = classes[i];
            const float vi = (bi.z - bi.x) * (bi.w - bi.y);
            for (int j = i + 1 + threadIdx.x; j < n; j += kNpNmsThreads) {
                if (!s_keep[j] || classes[j] != ci) continue;
                const float4 bj = boxes[j];
                const float h = fmaxf(fminf(bi.z, bj.z) - fmaxf(bi.x, bj.x), 0.f);
                const float w = fmaxf(fminf(bi.w, bj.w) - fmaxf(bi.y, bj.y), 0.f);
                const float inter = h * w;
                const float vj = (bj.z - bj.x) * (bj.w - bj.y);
                const float jac = __fdiv_rn(inter, (vi + vj) - inter);      // :197-200
                if (!(jac < thr)) s_keep[j] = 0;                     // keep_overlap = overlap < thr or other class
            }
        }
        __syncthreads();
    }
    for (int j = threadIdx.x; j < n; j += kNpNmsThreads) keep_out[j] = s_keep[j];
}

}  // namespace ronk

using namespace ronk;

extern "C" int ronk_np_select_mask(const float* pred, long long n, int C, int use_threshold, float threshold,
                                   uint8_t* out_mask, void* stream) {
    RONK_REQUIRE(n >= 0 && C >= 2 && (n == 0 || (pred && out_mask)), RONK_EINVAL, "ronk_np_select_mask: bad argument");
    const long long total = use_threshold ? n * (C - 1) : n;
    RONK_REQUIRE(total < (1ll << 31), RONK_ELIMIT, "ronk_np_select_mask: more than 2^31 candidates");
    if (total == 0) return RONK_OK;
    np_select_mask_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pred, n, C, use_threshold,
                                                                                         threshold, out_mask);
    RONK_LAUNCHED();
    return RONK_OK;
}

extern "C" int ronk_np_select_gather(const float* pred, const float* boxes, int C, int use_threshold, const int32_t* idx,
                                     int m, int64_t* out_classes, float* out_scores, float* out_boxes, void* stream) {
    RONK_REQUIRE(m >= 0 && C >= 2 && (m == 0 || (pred && boxes && idx && out_classes && out_scores && out_boxes)),
                 RONK_EINVAL, "ronk_np_select_gather: bad argument");
    if (m == 0) return RONK_OK;
    RONK_REQUIRE(((uintptr_t)boxes % 16) == 0 && ((uintptr_t)out_boxes % 16) == 0, RONK_EINVAL,
                 "ronk_np_select_gather: box pointers must be 16-byte aligned");
    np_select_gather_kernel<<<(unsigned)((m + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        pred, (const float4*)boxes, C, use_threshold, idx, m, (long long*)out_classes, out_scores, (float4*)out_boxes);
    RONK_LAUNCHED();
    return RONK_OK;
}

extern "C" int ronk_np_clip(const float* bbox_ref, const float* boxes, long long n, float* out_boxes, void* stream) {
    RONK_REQUIRE(bbox_ref && n >= 0 && (n == 0 || (boxes && out_boxes)), RONK_EINVAL, "ronk_np_clip: bad argument");
    if (n == 0) return RONK_OK;
    RONK_REQUIRE(((uintptr_t)boxes % 16) == 0 && ((uintptr_t)out_boxes % 16) == 0, RONK_EINVAL,
                 "ronk_np_clip: pointers must be 16-byte aligned");
    np_clip_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)boxes, n, make_float4(bbox_ref[0], bbox_ref[1], bbox_ref[2], bbox_ref[3]), (float4*)out_boxes);
    RONK_LAUNCHED();
    return RONK_OK;
}

extern "C" int ronk_np_nms(const int64_t* classes, const float* boxes, int n, float nms_threshold, uint8_t* out_keep,
                           void* stream) {
    RONK_REQUIRE(n >= 0 && (n == 0 || (classes && boxes && out_keep)), RONK_EINVAL, "ronk_np_nms: bad argument");
    if (n == 0) return RONK_OK;
    RONK_REQUIRE(n <= 200 * 1024, RONK_ELIMIT, "ronk_np_nms: at most 204800 boxes (keep flags live in shared memory)");
    RONK_REQUIRE(((uintptr_t)boxes % 16) == 0, RONK_EINVAL, "ronk_np_nms: boxes must be 16-byte aligned");
    const size_t smem = ((size_t)n + 15) & ~(size_t)15;
    if (smem > 48 * 1024)
        RONK_CUDA(cudaFuncSetAttribute(np_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    np_nms_kernel<<<1, kNpNmsThreads, smem, (cudaStream_t)stream>>>((const long long*)classes, (const float4*)boxes, n,
                                                                   nms_threshold, out_keep);
    RONK_LAUNCHED();
    return RONK_OK;
}
