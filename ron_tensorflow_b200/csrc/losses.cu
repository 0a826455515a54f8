// RON loss example masks + smooth-L1: the step right after the encode.
// Reference: nets/ron_vgg_320.py:686-740 (positive / negative masks, random negative sampling for the
// objectness loss and for the objectness-gated class loss), :760-764 (localisation loss) and
// nets/custom_layers.py:31-50 (modified_smooth_l1).
//   loss_fused_kernel   ONE cooperative launch (a co-resident grid), 4 elements per thread
//                       and step with 16-byte accesses: pass 1 counts the four example classes (positives,
//                       negatives, class positives, class negatives; warp reduction, one atomic per block and
//                       counter), grid.sync(), pass 2 derives the
//                       selection probabilities in the reference's float32 / int32 steps, writes the four masks
//                       (the second read of the labels comes from L2) and, when the localisations are given,
//                       accumulates the localisation term of the class positives; the CTA that finishes last
//                       writes counts + loss and restores the workspace to zero.  The two tf.random_uniform
//                       draws are inputs;
//   smooth_l1_kernel    element-wise modified_smooth_l1, one rounding per op;
//   loc_loss_kernel / loc_loss_finish_kernel   beta * mean over the class positives of the row sums,
//                       accumulated in double (the reference's reduction order is unspecified).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace ronk {

// tfe.safe_divide(cast(min(int32(ratio * n_pos), int32(n_neg)), f32), n_neg)   (ron_vgg_320.py:700-705)
__device__ __forceinline__ float select_prob(float ratio, unsigned n_pos, unsigned n_neg) {
    const float fp = (float)n_pos, fn = (float)n_neg;          // float32 sums of 0/1 masks
    const int want = (int)(ratio * fp);
    const int sel = min(want, (int)fn);
    return fn > 0.f ? (float)sel / fn : 0.f;
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

struct LossParams {
    const long long* gclasses;
    const float* objness;
    const float* rand_obj;
    const float* rand_cls;
    long long n;
    float obj_thr, ratio;
    uint8_t* final_obj;
    int* obj_label;
    uint8_t* cls_pos;
    uint8_t* final_cls;
    float* out_counts;          // [4] or NULL
    const float4* loc;          // [n] or NULL: no localisation term
    const float4* gloc;
    SmoothL1 f;
    float beta;
    float* out_loss;            // [1] or NULL
    unsigned* counts;           // ws: [4] example counts, [4] = CTAs finished      (zero between calls)
    double* acc;                // ws: [0] sum of row losses, [1] class positives   (zero between calls)
};

// VEC: every array is 16-byte aligned (4-byte for the uint8 masks): a thread handles 4 consecutive elements
// per step with 16-byte loads / stores, so that all of a thread's loads of a pass are in flight together; the
// (n % 4) tail and the unaligned case go through the scalar path.
template <bool VEC>
__global__ void __launch_bounds__(256, 4)
loss_fused_kernel(const __grid_constant__ LossParams p) {
    __shared__ unsigned s_c[4];
    __shared__ double s_sum[8];
    __shared__ int s_last;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long quads = VEC ? (p.n >> 2) : 0;
    if (threadIdx.x < 4) s_c[threadIdx.x] = 0u;
    __syncthreads();
    // ---- pass 1: counts
    {
        unsigned c0 = 0, c1 = 0, c2 = 0, c3 = 0;
        auto count = [&](long long g, float o) {
            const bool om = o > p.obj_thr, pos = g > 0, neg = g == 0;
            c0 += pos; c1 += neg; c2 += pos && om; c3 += neg && om;
        };
        for (long long q = gtid; q < quads; q += stride) {
            const longlong2 ga = reinterpret_cast<const longlong2*>(p.gclasses)[2 * q];
            const longlong2 gb = reinterpret_cast<const longlong2*>(p.gclasses)[2 * q + 1];
            const float4 o = reinterpret_cast<const float4*>(p.objness)[q];
            count(ga.x, o.x); count(ga.y, o.y); count(gb.x, o.z); count(gb.y, o.w);
        }
        for (long long i = 4 * quads + gtid; i < p.n; i += stride) count(p.gclasses[i], p.objness[i]);
        c0 = __reduce_add_sync(full, c0); c1 = __reduce_add_sync(full, c1);
        c2 = __reduce_add_sync(full, c2); c3 = __reduce_add_sync(full, c3);
        if (lane == 0) { atomicAdd(&s_c[0], c0); atomicAdd(&s_c[1], c1); atomicAdd(&s_c[2], c2); atomicAdd(&s_c[3], c3); }
        __syncthreads();
        if (threadIdx.x < 4 && s_c[threadIdx.x]) atomicAdd(p.counts + threadIdx.x, s_c[threadIdx.x]);
    }
    cg::this_grid().sync();
    // ---- pass 2: masks (+ localisation term of the class positives)
    unsigned cnt[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) cnt[k] = __ldcg(p.counts + k);
    const float p_obj = select_prob(p.ratio, cnt[0], cnt[1]);
    const float p_cls = select_prob(p.ratio, cnt[2], cnt[3]);
    double sum = 0.;
    // returns {final_obj, obj_label, cls_pos, final_cls} bits 0..3
    auto masks = [&](long long i, long long g, float o, float r1, float r2) -> unsigned {
        const bool pos = g > 0, neg = g == 0, om = o > p.obj_thr;
        const bool cp = pos && om, cn = om && neg;
        if (p.loc && cp) {
            const float4 a = p.loc[i], b = p.gloc[i];
            sum += (double)p.f(a.x, b.x) + (double)p.f(a.y, b.y) + (double)p.f(a.z, b.z) + (double)p.f(a.w, b.w);
        }
        return (((neg && r1 < p_obj) || pos) ? 1u : 0u) | (pos ? 2u : 0u) | (cp ? 4u : 0u) |
               (((cn && r2 < p_cls) || cp) ? 8u : 0u);
    };
    for (long long q = gtid; q < quads; q += stride) {
        const longlong2 ga = reinterpret_cast<const longlong2*>(p.gclasses)[2 * q];
        const longlong2 gb = reinterpret_cast<const longlong2*>(p.gclasses)[2 * q + 1];
        const float4 o = reinterpret_cast<const float4*>(p.objness)[q];
        const float4 r1 = reinterpret_cast<const float4*>(p.rand_obj)[q];
        const float4 r2 = reinterpret_cast<const float4*>(p.rand_cls)[q];
        const unsigned m0 = masks(4 * q, ga.x, o.x, r1.x, r2.x), m1 = masks(4 * q + 1, ga.y, o.y, r1.y, r2.y);
        const unsigned m2 = masks(4 * q + 2, gb.x, o.z, r1.z, r2.z), m3 = masks(4 * q + 3, gb.y, o.w, r1.w, r2.w);
        auto pack = [&](int bit) {
            return make_uchar4((m0 >> bit) & 1u, (m1 >> bit) & 1u, (m2 >> bit) & 1u, (m3 >> bit) & 1u);
        };
        reinterpret_cast<uchar4*>(p.final_obj)[q] = pack(0);
        reinterpret_cast<int4*>(p.obj_label)[q] = make_int4((m0 >> 1) & 1, (m1 >> 1) & 1, (m2 >> 1) & 1, (m3 >> 1) & 1);
        reinterpret_cast<uchar4*>(p.cls_pos)[q] = pack(2);
        reinterpret_cast<uchar4*>(p.final_cls)[q] = pack(3);
    }
    for (long long i = 4 * quads + gtid; i < p.n; i += stride) {
        const unsigned m = masks(i, p.gclasses[i], p.objness[i], p.rand_obj[i], p.rand_cls[i]);
        p.final_obj[i] = m & 1u;
        p.obj_label[i] = (m >> 1) & 1u;
        p.cls_pos[i] = (m >> 2) & 1u;
        p.final_cls[i] = (m >> 3) & 1u;
    }
    if (p.loc) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(full, sum, o);
        if (lane == 0) s_sum[warp] = sum;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (p.loc) {
            for (int w = 1; w < 8; ++w) sum += s_sum[w];
            if (sum != 0.) atomicAdd(p.acc, sum);
        }
        __threadfence();
        s_last = atomicAdd(p.counts + 4, 1u) == gridDim.x - 1u;
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        if (p.out_counts)
            for (int k = 0; k < 4; ++k) p.out_counts[k] = (float)cnt[k];
        if (p.out_loss) {
            // tf.cond(n_cls_positives > 0, beta * reduce_mean(...), 0)   (ron_vgg_320.py:764)
            const double tot = __ldcg(p.acc);
            p.out_loss[0] = (p.loc && cnt[2] > 0u) ? (float)((double)p.beta * (tot / (double)cnt[2])) : 0.f;
        }
        for (int k = 0; k < 5; ++k) p.counts[k] = 0u;
        p.acc[0] = 0.;
    }
}

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

// d SmoothL1 / d pred of one element (the comparison is a constant for the gradient, like tf.cast(tf.less(...)));
// tf.abs has gradient sign(x), 0 at 0
__device__ __forceinline__ float smooth_l1_grad(const SmoothL1& f, float pred, float target) {
    const float x = f.in_w * (pred - target);
    const float d = fabsf(x) < f.thr ? (x + x) * f.half_s2 : (x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f));
    return f.out_w * d * f.in_w;
}

__global__ void __launch_bounds__(256)
smooth_l1_backward_kernel(const float* __restrict__ pred, const float* __restrict__ target, const float* __restrict__ grad_out,
                          long long count, SmoothL1 f, float* __restrict__ grad_pred) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) grad_pred[i] = grad_out[i] * smooth_l1_grad(f, pred[i], target[i]);
}

__global__ void __launch_bounds__(256)
mask_count_kernel(const uint8_t* __restrict__ mask, long long n, double* __restrict__ acc) {
    unsigned cnt = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) cnt += mask[i] ? 1u : 0u;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(acc, (double)cnt);
}

// d (beta * mean over the masked rows of the row sums) / d localisations, times the incoming scalar gradient
__global__ void __launch_bounds__(256)
loc_loss_backward_kernel(const float4* __restrict__ loc, const float4* __restrict__ gloc, const uint8_t* __restrict__ mask,
                         long long n, SmoothL1 f, float beta, const double* __restrict__ acc, const float* __restrict__ grad_out,
                         float4* __restrict__ grad_loc) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (mask[i] && acc[0] > 0.) {
        const float s = grad_out[0] * (float)((double)beta / acc[0]);
        const float4 a = loc[i], b = gloc[i];
        g = make_float4(s * smooth_l1_grad(f, a.x, b.x), s * smooth_l1_grad(f, a.y, b.y), s * smooth_l1_grad(f, a.z, b.z),
                        s * smooth_l1_grad(f, a.w, b.w));
    }
    grad_loc[i] = g;
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

extern "C" size_t ronk_loss_workspace_bytes(void) { return 64; }

extern "C" int ronk_loss_workspace_init(void* ws, void* stream) {
    RONK_REQUIRE(ws != nullptr, RONK_EINVAL, "ronk_loss_workspace_init: NULL workspace");
    RONK_CUDA(cudaMemsetAsync(ws, 0, 64, (cudaStream_t)stream));
    return RONK_OK;
}

extern "C" int ronk_loss_masks(const int64_t* gclasses, const float* objness_pred, const float* rand_objness,
                               const float* rand_cls, long long n, float objness_threshold, float negative_ratio,
                               uint8_t* out_final_objness, int32_t* out_objness_label, uint8_t* out_cls_positive,
                               uint8_t* out_final_cls, float* out_counts, const float* localisations,
                               const float* glocalisations, double sigma, float beta, float* out_loss, void* ws,
                               void* stream) {
    RONK_REQUIRE(n >= 0 && ws, RONK_EINVAL, "ronk_loss_masks: bad argument");
    RONK_REQUIRE(n < (1ll << 24), RONK_ELIMIT, "ronk_loss_masks: the reference counts in float32: n must stay below 2^24");
    RONK_REQUIRE(n == 0 || (gclasses && objness_pred && rand_objness && rand_cls && out_final_objness && out_objness_label &&
                            out_cls_positive && out_final_cls),
                 RONK_EINVAL, "ronk_loss_masks: NULL pointer argument");
    RONK_REQUIRE((localisations == nullptr) == (glocalisations == nullptr), RONK_EINVAL,
                 "ronk_loss_masks: localisations and glocalisations go together");
    RONK_REQUIRE(!localisations || (sigma > 0. && ((uintptr_t)localisations % 16) == 0 && ((uintptr_t)glocalisations % 16) == 0),
                 RONK_EINVAL, "ronk_loss_masks: sigma must be positive and the box pointers 16-byte aligned");
    RONK_REQUIRE(((uintptr_t)ws % 8) == 0, RONK_EINVAL, "ronk_loss_masks: ws must be 8-byte aligned");
    LossParams p;
    p.gclasses = (const long long*)gclasses;
    p.objness = objness_pred;
    p.rand_obj = rand_objness;
    p.rand_cls = rand_cls;
    p.n = n;
    p.obj_thr = objness_threshold;
    p.ratio = negative_ratio;
    p.final_obj = out_final_objness;
    p.obj_label = out_objness_label;
    p.cls_pos = out_cls_positive;
    p.final_cls = out_final_cls;
    p.out_counts = out_counts;
    p.loc = (const float4*)localisations;
    p.gloc = (const float4*)glocalisations;
    p.f = make_smooth_l1(1.f, 1.f, localisations ? sigma : 1.);
    p.beta = beta;
    p.out_loss = out_loss;
    p.counts = (unsigned*)ws;
    p.acc = (double*)((char*)ws + 32);
    const uintptr_t all16 = (uintptr_t)gclasses | (uintptr_t)objness_pred | (uintptr_t)rand_objness | (uintptr_t)rand_cls |
                            (uintptr_t)out_objness_label;
    const uintptr_t all4 = (uintptr_t)out_final_objness | (uintptr_t)out_cls_positive | (uintptr_t)out_final_cls;
    const bool vec = (all16 % 16) == 0 && (all4 % 4) == 0;
    const void* kern = vec ? (const void*)loss_fused_kernel<true> : (const void*)loss_fused_kernel<false>;
    int dev = 0, sms = 0, per_sm = 0;
    RONK_CUDA(cudaGetDevice(&dev));
    RONK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (vec) RONK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, loss_fused_kernel<true>, 256, 0));
    else RONK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, loss_fused_kernel<false>, 256, 0));
    RONK_REQUIRE(per_sm >= 1, RONK_ECUDA, "ronk_loss_masks: the kernel does not fit on an SM");
    long long blocks = ((vec ? (n + 3) / 4 : n) + 255) / 256;
    const long long resident = (long long)sms * per_sm;        // grid.sync() needs every CTA resident
    if (blocks > resident) blocks = resident;
    if (blocks < 1) blocks = 1;
    void* args[] = {(void*)&p};
    RONK_CUDA(cudaLaunchCooperativeKernel(kern, dim3((unsigned)blocks), dim3(256), args, 0, (cudaStream_t)stream));
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

extern "C" int ronk_smooth_l1_backward(const float* pred, const float* target, const float* grad_out, long long count,
                                       float inside_weight, float outside_weight, double sigma, float* grad_pred, void* stream) {
    RONK_REQUIRE(count >= 0 && sigma > 0. && (count == 0 || (pred && target && grad_out && grad_pred)), RONK_EINVAL,
                 "ronk_smooth_l1_backward: bad argument");
    if (count == 0) return RONK_OK;
    smooth_l1_backward_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        pred, target, grad_out, count, make_smooth_l1(inside_weight, outside_weight, sigma), grad_pred);
    RONK_LAUNCHED();
    return RONK_OK;
}

extern "C" int ronk_localization_loss_backward(const float* localisations, const float* glocalisations,
                                               const uint8_t* cls_positive, long long n, double sigma, float beta,
                                               const float* grad_out, float* grad_localisations, void* ws, void* stream) {
    RONK_REQUIRE(n >= 0 && sigma > 0. && grad_out && ws && (n == 0 || (localisations && glocalisations && cls_positive && grad_localisations)),
                 RONK_EINVAL, "ronk_localization_loss_backward: bad argument");
    RONK_REQUIRE(((uintptr_t)localisations % 16) == 0 && ((uintptr_t)glocalisations % 16) == 0 &&
                 ((uintptr_t)grad_localisations % 16) == 0 && ((uintptr_t)ws % 8) == 0, RONK_EINVAL,
                 "ronk_localization_loss_backward: box pointers must be 16-byte aligned, ws 8-byte aligned");
    if (n == 0) return RONK_OK;
    cudaStream_t st = (cudaStream_t)stream;
    double* acc = (double*)((char*)ws + 48);
    RONK_CUDA(cudaMemsetAsync(acc, 0, 16, st));
    const unsigned blocks = (unsigned)((n + 255) / 256);
    mask_count_kernel<<<blocks < 1184u ? blocks : 1184u, 256, 0, st>>>(cls_positive, n, acc);
    RONK_LAUNCHED();
    loc_loss_backward_kernel<<<blocks, 256, 0, st>>>((const float4*)localisations, (const float4*)glocalisations, cls_positive, n,
                                                     make_smooth_l1(1.f, 1.f, sigma), beta, acc, grad_out,
                                                     (float4*)grad_localisations);
    RONK_LAUNCHED();
    RONK_CUDA(cudaMemsetAsync(acc, 0, 16, st));
    return RONK_OK;
}

extern "C" int ronk_localization_loss(const float* localisations, const float* glocalisations, const uint8_t* cls_positive,
                                      long long n, double sigma, float beta, float* out_loss, void* ws, void* stream) {
    RONK_REQUIRE(n >= 0 && sigma > 0. && out_loss && ws && (n == 0 || (localisations && glocalisations && cls_positive)),
                 RONK_EINVAL, "ronk_localization_loss: bad argument");
    RONK_REQUIRE(((uintptr_t)localisations % 16) == 0 && ((uintptr_t)glocalisations % 16) == 0 && ((uintptr_t)ws % 8) == 0,
                 RONK_EINVAL, "ronk_localization_loss: box pointers must be 16-byte aligned, ws 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    double* acc = (double*)((char*)ws + 48);             // bytes 48..63: not shared with ronk_loss_masks
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
