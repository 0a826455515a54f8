// RON loss example masks + smooth-L1: the step right after the encode.
// Reference: nets/ron_vgg_320.py:686-740 (positive / negative masks, random negative sampling for the
// objectness loss and for the objectness-gated class loss), :760-764 (localisation loss) and
// nets/custom_layers.py:31-50 (modified_smooth_l1).
//   loss_count_kernel   the four example counts (positives, negatives, class positives, class negatives):
//                       ballot + popc per warp, one atomic per block and counter;
//   loss_mask_kernel    selection probabilities from the counts in the reference's float32 / int32 steps,
//                       then the four masks; the two tf.random_uniform draws are inputs;
//   smooth_l1_kernel    element-wise modified_smooth_l1, one rounding per op;
//   loc_loss_kernel / loc_loss_finish_kernel   beta * mean over the class positives of the row sums,
//                       accumulated in double (the reference's reduction order is unspecified).
#include "common.cuh"

namespace ronk {

__global__ void __launch_bounds__(256)
loss_count_kernel(const long long* __restrict__ gclasses, const float* __restrict__ objness, long long n, float obj_thr,
                  unsigned* __restrict__ counts) {
    __shared__ unsigned s_c[4];
    if (threadIdx.x < 4) s_c[threadIdx.x] = 0u;
    __syncthreads();
    const unsigned full = 0xffffffffu;
    unsigned c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    // whole warps iterate together (the ballots need every lane): round the bound up to a warp
    for (long long i0 = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < n; i0 += stride) {
        const long long i = i0 + (threadIdx.x & 31);
        const bool in = i < n;
        const long long g = in ? gclasses[i] : -1;
        const bool om = in && objness[i] > obj_thr;
        const bool pos = g > 0, neg = g == 0;
        c0 += __popc(__ballot_sync(full, pos));
        c1 += __popc(__ballot_sync(full, neg));
        c2 += __popc(__ballot_sync(full, pos && om));
        c3 += __popc(__ballot_sync(full, neg && om));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_c[0], c0); atomicAdd(&s_c[1], c1); atomicAdd(&s_c[2], c2); atomicAdd(&s_c[3], c3);
    }
    __syncthreads();
    if (threadIdx.x < 4 && s_c[threadIdx.x]) atomicAdd(counts + threadIdx.x, s_c[threadIdx.x]);
}

// tfe.safe_divide(cast(min(int32(ratio * n_pos), int32(n_neg)), f32), n_neg)   (ron_vgg_320.py:700-705)
__device__ __forceinline__ float select_prob(float ratio, unsigned n_pos, unsigned n_neg) {
    const float fp = (float)n_pos, fn = (float)n_neg;          // float32 sums of 0/1 masks
    const int want = (int)(ratio * fp);
    const int sel = min(want, (int)fn);
    return fn > 0.f ? (float)sel / fn : 0.f;
}

__global__ void __launch_bounds__(256)
loss_mask_kernel(const long long* __restrict__ gclasses, const float* __restrict__ objness,
                 const float* __restrict__ rand_obj, const float* __restrict__ rand_cls, long long n, float obj_thr,
                 float ratio, const unsigned* __restrict__ counts, uint8_t* __restrict__ final_obj,
                 int* __restrict__ obj_label, uint8_t* __restrict__ cls_pos, uint8_t* __restrict__ final_cls,
                 float* __restrict__ out_counts) {
    const float p_obj = select_prob(ratio, counts[0], counts[1]);
    const float p_cls = select_prob(ratio, counts[2], counts[3]);
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && out_counts) {
        for (int k = 0; k < 4; ++k) out_counts[k] = (float)counts[k];
    }
    if (i >= n) return;
    const long long g = gclasses[i];
    const bool pos = g > 0, neg = g == 0;
    const bool om = objness[i] > obj_thr;
    const bool cp = pos && om, cn = om && neg;
    final_obj[i] = ((neg && rand_obj[i] < p_obj) || pos) ? 1 : 0;
    obj_label[i] = pos ? 1 : 0;
    cls_pos[i] = cp ? 1 : 0;
    final_cls[i] = ((cn && rand_cls[i] < p_cls) || cp) ? 1 : 0;
}

struct SmoothL1 {
    float in_w, out_w, thr, half_s2, off;        // 1/sigma^2, 0.5*sigma^2, 0.5/sigma^2 (double on the host, then f32)
    __device__ __forceinline__ float operator()(float pred, float target) const {
        const float x = in_w * (pred - target);
        const float sign = fabsf(x) < thr ? 1.f : 0.f;
        const float o1 = (x * x) * half_s2;
        const float o2 = fabsf(x) - off;
        return out_w * (o1 * sign + o2 * fabsf(sign - 1.f));
    }
};

__global__ void __launch_bounds__(256)
smooth_l1_kernel(const float* __restrict__ pred, const float* __restrict__ target, long long count, SmoothL1 f,
                 float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = f(pred[i], target[i]);
}

__global__ void __launch_bounds__(256)
loc_loss_kernel(const float4* __restrict__ loc, const float4* __restrict__ gloc, const uint8_t* __restrict__ mask,
                long long n, SmoothL1 f, double* __restrict__ acc /*[2]: sum, count*/) {
    __shared__ double s_sum[8];
    __shared__ unsigned s_cnt[8];
    double sum = 0.;
    unsigned cnt = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (!mask[i]) continue;
        const float4 a = loc[i], b = gloc[i];
        sum += (double)f(a.x, b.x) + (double)f(a.y, b.y) + (double)f(a.z, b.z) + (double)f(a.w, b.w);
        ++cnt;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = sum; s_cnt[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { sum += s_sum[w]; cnt += s_cnt[w]; }
        if (cnt) { atomicAdd(acc, sum); atomicAdd(acc + 1, (double)cnt); }
    }
}

__global__ void loc_loss_finish_kernel(const double* __restrict__ acc, float beta, float* __restrict__ out) {
    // tf.cond(n_cls_positives > 0, beta * reduce_mean(...), 0)   (ron_vgg_320.py:764)
    out[0] = acc[1] > 0. ? (float)((double)beta * (acc[0] / acc[1])) : 0.f;
}

static SmoothL1 make_smooth_l1(float inside_w, float outside_w, double sigma) {
    const double s2 = sigma * sigma;
    SmoothL1 f;
    f.in_w = inside_w;
    f.out_w = outside_w;
    f.thr = (float)(1.0 / s2);
    f.half_s2 = (float)(0.5 * s2);
    f.off = (float)(0.5 / s2);
    return f;
}

}  // namespace ronk

using namespace ronk;

extern "C" size_t ronk_loss_workspace_bytes(void) { return 32; }

extern "C" int ronk_loss_masks(const int64_t* gclasses, const float* objness_pred, const float* rand_objness,
                               const float* rand_cls, long long n, float objness_threshold, float negative_ratio,
                               uint8_t* out_final_objness, int32_t* out_objness_label, uint8_t* out_cls_positive,
                               uint8_t* out_final_cls, float* out_counts, void* ws, void* stream) {
    RONK_REQUIRE(n >= 0 && ws, RONK_EINVAL, "ronk_loss_masks: bad argument");
    RONK_REQUIRE(n < (1ll << 24), RONK_ELIMIT, "ronk_loss_masks: the reference counts in float32: n must stay below 2^24");
    RONK_REQUIRE(n == 0 || (gclasses && objness_pred && rand_objness && rand_cls && out_final_objness && out_objness_label &&
                            out_cls_positive && out_final_cls),
                 RONK_EINVAL, "ronk_loss_masks: NULL pointer argument");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned* counts = (unsigned*)ws;
    RONK_CUDA(cudaMemsetAsync(counts, 0, 16, st));
    if (n > 0) {
        const unsigned blocks = (unsigned)((n + 255) / 256);
        loss_count_kernel<<<blocks < 1184u ? blocks : 1184u, 256, 0, st>>>((const long long*)gclasses, objness_pred, n,
                                                                           objness_threshold, counts);
        RONK_LAUNCHED();
    }
    loss_mask_kernel<<<n > 0 ? (unsigned)((n + 255) / 256) : 1u, 256, 0, st>>>(
        (const long long*)gclasses, objness_pred, rand_objness, rand_cls, n, objness_threshold, negative_ratio, counts,
        out_final_objness, out_objness_label, out_cls_positive, out_final_cls, out_counts);
    RONK_LAUNCHED();
    return RONK_OK;
}

extern "C" int ronk_smooth_l1(const float* pred, const float* target, long long count, float inside_weight,
                              float outside_weight, double sigma, float* out, void* stream) {
    RONK_REQUIRE(count >= 0 && sigma > 0. && (count == 0 || (pred && target && out)), RONK_EINVAL,
                 "ronk_smooth_l1: bad argument");
    if (count == 0) return RONK_OK;
    smooth_l1_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        pred, target, count, make_smooth_l1(inside_weight, outside_weight, sigma), out);
    RONK_LAUNCHED();
    return RONK_OK;
}

extern "C" int ronk_localization_loss(const float* localisations, const float* glocalisations, const uint8_t* cls_positive,
                                      long long n, double sigma, float beta, float* out_loss, void* ws, void* stream) {
    RONK_REQUIRE(n >= 0 && sigma > 0. && out_loss && ws && (n == 0 || (localisations && glocalisations && cls_positive)),
                 RONK_EINVAL, "ronk_localization_loss: bad argument");
    RONK_REQUIRE(((uintptr_t)localisations % 16) == 0 && ((uintptr_t)glocalisations % 16) == 0 && ((uintptr_t)ws % 8) == 0,
                 RONK_EINVAL, "ronk_localization_loss: box pointers must be 16-byte aligned, ws 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    double* acc = (double*)ws + 2;                      // bytes 16..31 of the workspace
    RONK_CUDA(cudaMemsetAsync(acc, 0, 16, st));
    if (n > 0) {
        const unsigned blocks = (unsigned)((n + 255) / 256);
        loc_loss_kernel<<<blocks < 1184u ? blocks : 1184u, 256, 0, st>>>((const float4*)localisations,
                                                                         (const float4*)glocalisations, cls_positive, n,
                                                                         make_smooth_l1(1.f, 1.f, sigma), acc);
        RONK_LAUNCHED();
    }
    loc_loss_finish_kernel<<<1, 1, 0, st>>>(acc, beta, out_loss);
    RONK_LAUNCHED();
    return RONK_OK;
}
