// K1: fused all-layer GT->anchor matching + target encoding.
//
// Reference: nets/ssd_common.py:27-47 (iou_matrix), :49-75 (do_dual_max_match), :77-147
// (tf_ssd_bboxes_encode_layer), nets/ron_vgg_320.py:686,708 (objectness label).
// Spec: SURVEY.md Appendix A.3/A.4.  Results are bit-exact against oracle/ron_oracle.py.
//
// Two launches per batch:
//   match_encode_kernel  one CTA = one tile of consecutive anchors (kept in registers) x a
//     group of images.  Per image the CTA culls the GT list against the tile's bounding box
//     (order preserving, so "first GT wins" ties need no extra compare), evaluates IoU only
//     where the intersection is positive (everything else is exactly 0 in float32), keeps
//     the per-anchor (max, first argmax) in registers, reduces the per-GT (max, lowest
//     anchor) as a packed u64 with REDUX + shared/global atomicMax, then labels, encodes and
//     writes labels/loc/scores fully coalesced.  No [G,N] matrix ever exists.
//   match_force_kernel   per image: reads the per-GT best anchors, applies "lowest GT index
//     claims the anchor", rewrites those (<= G) anchors, and re-zeroes the workspace.
#include <math_constants.h>

#include "common.cuh"

namespace ronk {

constexpr int kEncThreads = 256;
constexpr int kEncApt = 2;   // anchors per thread

struct EncodeParams {
    const float4* cor;
    const float4* mcor;
    const float4* enc;
    const uint8_t* inside;
    int N;
    const float4* gt_boxes;
    const long long* gt_labels;
    const int* gt_counts;
    int B, Gmax, gcap;
    float high, low;
    float ps0, ps1, ps2, ps3;
    int ignore_between, gt_max_first;
    long long* out_labels;
    float4* out_loc;
    float* out_scores;
    int* out_matched;
    int* out_obj;
    u64* ws_keys;
    unsigned* ws_claimed;
};

// nets/ssd_common.py:130-144: (cx, cy, w, h) ordering, two true divisions per term.
// encode divides y by ps0, x by ps1, h by ps2, w by ps3 (Appendix A.5 note).
__device__ __forceinline__ float4 encode_loc(float4 gb, float4 e, const EncodeParams& p) {
    float gcy = (gb.z + gb.x) / 2.f;
    float gcx = (gb.w + gb.y) / 2.f;
    float gh = gb.z - gb.x;
    float gw = gb.w - gb.y;
    float t_cy = ((gcy - e.x) / e.z) / p.ps0;
    float t_cx = ((gcx - e.y) / e.w) / p.ps1;
    float t_h = log_cr(gh / e.z) / p.ps2;
    float t_w = log_cr(gw / e.w) / p.ps3;
    return make_float4(t_cx, t_cy, t_w, t_h);
}

template <int APT>
__global__ void __launch_bounds__(kEncThreads, 4)
match_encode_kernel(const __grid_constant__ EncodeParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    float4* s_box = reinterpret_cast<float4*>(smem);              // [gcap] GT corners
    u64* s_best = reinterpret_cast<u64*>(s_box + p.gcap);         // [gcap] per-GT (iou bits, ~anchor) of this tile
    float* s_area = reinterpret_cast<float*>(s_best + p.gcap);    // [gcap]

    const unsigned full = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int warp_n0 = blockIdx.x * (kEncThreads * APT) + warp * (32 * APT);

    // ---- this thread's anchors: corners + area in registers for the whole image loop.  The
    // match table already holds an empty box (+inf,+inf,-inf,-inf) for anchors outside the
    // border mask: their overlap with everything is exactly 0 (ssd_common.py:118).
    float4 a[APT];
    float area[APT];
    float wy0 = CUDART_INF_F, wx0 = CUDART_INF_F, wy1 = -CUDART_INF_F, wx1 = -CUDART_INF_F;
#pragma unroll
    for (int j = 0; j < APT; ++j) {
        int n = warp_n0 + j * 32 + lane;
        a[j] = (n < p.N) ? p.mcor[n] : make_float4(CUDART_INF_F, CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
        const bool in = a[j].x != CUDART_INF_F;          // false only for the empty box
        area[j] = in ? (a[j].w - a[j].y) * (a[j].z - a[j].x) : 0.f;
        wy0 = fminf(wy0, a[j].x);
        wx0 = fminf(wx0, a[j].y);
        wy1 = fmaxf(wy1, a[j].z);
        wx1 = fmaxf(wx1, a[j].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        wy0 = fminf(wy0, __shfl_xor_sync(full, wy0, o));
        wx0 = fminf(wx0, __shfl_xor_sync(full, wx0, o));
        wy1 = fmaxf(wy1, __shfl_xor_sync(full, wy1, o));
        wx1 = fmaxf(wx1, __shfl_xor_sync(full, wx1, o));
    }

    // software prefetch of the next image's GT boxes (one per thread; G > 256 reads the rest late)
    int b = blockIdx.y;
    int G = 0;
    float4 gb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (b < p.B) {
        G = p.gt_counts[b];
        G = G < 0 ? 0 : (G > p.Gmax ? p.Gmax : G);
        if (tid < G) gb = p.gt_boxes[(size_t)b * p.Gmax + tid];
    }
    for (; b < p.B;) {
        const float4* gtb = p.gt_boxes + (size_t)b * p.Gmax;
        __syncthreads();                               // previous image is done with shared memory
        for (int g = tid; g < G; g += kEncThreads) {
            float4 v = (g == tid) ? gb : gtb[g];
            s_box[g] = v;
            s_area[g] = (v.w - v.y) * (v.z - v.x);
            s_best[g] = 0ull;
        }
        const int bn = b + gridDim.y;
        int Gn = 0;
        float4 gbn = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bn < p.B) {
            Gn = p.gt_counts[bn];
            Gn = Gn < 0 ? 0 : (Gn > p.Gmax ? p.Gmax : Gn);
            if (tid < Gn) gbn = p.gt_boxes[(size_t)bn * p.Gmax + tid];
        }
        __syncthreads();

        // ---- IoU sweep.  best/bestg: per-anchor running max and FIRST argmax over GT (strict
        // '>' over ascending g == tf.argmax first occurrence).  32 GT boxes are culled against
        // this warp's extent at once (one per lane, ballot); a pair can only have a positive
        // intersection if the GT overlaps the union extent (float subtraction is sign exact),
        // every skipped pair contributes exactly 0.
        float best[APT];
        int bestg[APT];
#pragma unroll
        for (int j = 0; j < APT; ++j) { best[j] = 0.f; bestg[j] = -1; }

        for (int g0 = 0; g0 < G; g0 += 32) {
            bool touch = false;
            if (g0 + lane < G) {
                float4 t = s_box[g0 + lane];
                touch = (fminf(t.z, wy1) > fmaxf(t.x, wy0)) && (fminf(t.w, wx1) > fmaxf(t.y, wx0));
            }
            unsigned todo = __ballot_sync(full, touch);
            while (todo) {
                const int g = g0 + __ffs(todo) - 1;
                todo &= todo - 1;
                const float4 t = s_box[g];
                const float ga = s_area[g];
                unsigned bits[APT];
#pragma unroll
                for (int j = 0; j < APT; ++j) {
                    // branch-free, exactly the reference's op order (ssd_common.py:34-47)
                    float h = fmaxf(fminf(t.z, a[j].z) - fmaxf(t.x, a[j].x), 0.f);
                    float w = fmaxf(fminf(t.w, a[j].w) - fmaxf(t.y, a[j].y), 0.f);
                    float inter = h * w;
                    float uni = (ga + area[j]) - inter;
                    float iou = (uni == 0.f) ? 0.f : inter / uni;
                    if (iou > best[j]) { best[j] = iou; bestg[j] = g; }
                    bits[j] = __float_as_uint(iou);
                }
                // per-GT (max, lowest anchor index): IoU >= 0 so float bits order as integers
                unsigned mybits = bits[0];
#pragma unroll
                for (int j = 1; j < APT; ++j) mybits = max(mybits, bits[j]);
                const unsigned m = __reduce_max_sync(full, mybits);
                if (m != 0u) {
                    unsigned cand = 0xffffffffu;
#pragma unroll
                    for (int j = APT - 1; j >= 0; --j)
                        if (bits[j] == m) cand = (unsigned)(j * 32 + lane);
                    const unsigned first = __reduce_min_sync(full, cand);
                    if (lane == 0) {
                        u64 key = ((u64)m << 32) | (u64)(0xffffffffu - ((unsigned)warp_n0 + first));
                        if (key > s_best[g]) atomicMax(&s_best[g], key);
                    }
                }
            }
        }
        __syncthreads();

        for (int g = tid; g < G; g += kEncThreads) {
            u64 v = s_best[g];
            if (v != 0ull) atomicMax(p.ws_keys + (size_t)b * p.Gmax + g, v);
        }

        // ---- label + encode + store (forced anchors are rewritten by match_force_kernel)
#pragma unroll
        for (int j = 0; j < APT; ++j) {
            int n = warp_n0 + j * 32 + lane;
            if (n >= p.N) continue;
            float mv = best[j];
            int a2g = bestg[j] < 0 ? 0 : bestg[j];
            bool less = mv < p.low;
            bool between = (mv < p.high) && (mv >= p.low);
            bool neg = p.ignore_between ? less : between;
            bool ign = p.ignore_between ? between : less;
            int mi = ign ? -2 : (neg ? -1 : a2g);
            if (G == 0) mi = -1;
            long long label = 0;
            float4 loc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (mi >= 0) {
                label = p.gt_labels[(size_t)b * p.Gmax + a2g];
                loc = encode_loc(s_box[a2g], p.enc[n], p);
                if (!p.gt_max_first) atomicOr(p.ws_claimed + (size_t)b * p.Gmax + a2g, 1u);
            } else if (mi < -1) {
                label = -1;
            }
            size_t o = (size_t)b * p.N + n;
            p.out_labels[o] = label;
            p.out_loc[o] = loc;
            p.out_scores[o] = mv;
            if (p.out_matched) p.out_matched[o] = mi;
            if (p.out_obj) p.out_obj[o] = label > 0 ? 1 : 0;
        }
        b = bn;
        G = Gn;
        gb = gbn;
    }
}

// Per image: g2a[g] = decoded per-GT best anchor (all-zero row -> anchor 0); the lowest GT
// index that claims an anchor wins (tf.argmax of the one-hot mask, ssd_common.py:74-75);
// score = overlap[g, n].  Also restores the workspace to zero for the next call.
__global__ void __launch_bounds__(128)
match_force_kernel(const __grid_constant__ EncodeParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    int* s_n = reinterpret_cast<int*>(smem);
    int* s_cl = s_n + p.Gmax;
    const int b = blockIdx.x;
    int G = p.gt_counts[b];
    G = G < 0 ? 0 : (G > p.Gmax ? p.Gmax : G);
    for (int g = threadIdx.x; g < p.Gmax; g += blockDim.x) {
        size_t o = (size_t)b * p.Gmax + g;
        u64 key = p.ws_keys[o];
        s_n[g] = key ? (int)(0xffffffffu - (unsigned)(key & 0xffffffffull)) : 0;
        s_cl[g] = p.gt_max_first ? 0 : (int)p.ws_claimed[o];
        p.ws_keys[o] = 0ull;
        p.ws_claimed[o] = 0u;
    }
    __syncthreads();
    for (int g0 = 0; g0 < G; g0 += blockDim.x) {
        const int g = g0 + threadIdx.x;
        const bool act = g < G;
        const int n = act ? s_n[g] : -1;
        // uniform trip count (no early exit): lanes must stay converged for the body below
        bool first = act && !s_cl[act ? g : 0];   // gt_max_first=False: a GT that already owns an anchor forces nothing
        const int lim = min(G, g0 + (int)blockDim.x);
        for (int g2 = 0; g2 < lim; ++g2) first = first && !(g2 < g && s_n[g2] == n && !s_cl[g2]);
        if (!first) continue;
        const float4 gb = p.gt_boxes[(size_t)b * p.Gmax + g];
        const float4 a = p.cor[n];
        const float4 e = p.enc[n];
        const bool in = p.inside[n] != 0;
        const long long label = p.gt_labels[(size_t)b * p.Gmax + g];
        float h = fmaxf(fminf(gb.z, a.z) - fmaxf(gb.x, a.x), 0.f);
        float w = fmaxf(fminf(gb.w, a.w) - fmaxf(gb.y, a.y), 0.f);
        float inter = h * w;
        float uni = ((gb.w - gb.y) * (gb.z - gb.x) + (a.w - a.y) * (a.z - a.x)) - inter;
        float iou = (uni == 0.f) ? 0.f : inter / uni;
        float ov = iou * (in ? 1.f : 0.f);
        size_t o = (size_t)b * p.N + n;
        p.out_labels[o] = label;
        p.out_loc[o] = encode_loc(gb, e, p);
        p.out_scores[o] = ov;
        if (p.out_matched) p.out_matched[o] = g;
        if (p.out_obj) p.out_obj[o] = label > 0 ? 1 : 0;
    }
}

__global__ void zero_u32_kernel(unsigned* p, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0u;
}

}  // namespace ronk

using namespace ronk;

extern "C" size_t ronk_encode_workspace_bytes(int B, int Gmax) {
    if (B < 1 || Gmax < 1) return 0;
    return (size_t)B * Gmax * (sizeof(u64) + sizeof(unsigned));
}

extern "C" int ronk_encode_workspace_init(void* ws, int B, int Gmax, void* stream) {
    RONK_REQUIRE(ws && B >= 1 && Gmax >= 1, RONK_EINVAL, "ronk_encode_workspace_init: bad argument");
    size_t words = ronk_encode_workspace_bytes(B, Gmax) / 4;
    zero_u32_kernel<<<(unsigned)((words + 255) / 256), 256, 0, (cudaStream_t)stream>>>((unsigned*)ws, words);
    RONK_LAUNCHED();
    return RONK_OK;
}

extern "C" int ronk_match_encode(const ronk_anchors_t* h, const float* gt_boxes, const int64_t* gt_labels,
                                 const int32_t* gt_counts, int B, int Gmax, float positive_threshold,
                                 float ignore_threshold, const float* ps, int match_flags,
                                 int64_t* out_labels, float* out_loc, float* out_scores,
                                 int32_t* out_matched, int32_t* out_objness, void* ws, void* stream) {
    RONK_REQUIRE(h != nullptr, RONK_EINVAL, "ronk_match_encode: NULL anchor handle");
    RONK_REQUIRE(gt_boxes && gt_labels && gt_counts && ps && out_labels && out_loc && out_scores && ws,
                 RONK_EINVAL, "ronk_match_encode: NULL pointer argument");
    RONK_REQUIRE(B >= 1 && Gmax >= 1, RONK_EINVAL, "ronk_match_encode: B and Gmax must be >= 1");
    RONK_REQUIRE(Gmax <= 1024, RONK_ELIMIT, "ronk_match_encode: Gmax > 1024 not supported");
    RONK_REQUIRE(((uintptr_t)gt_boxes % 16) == 0 && ((uintptr_t)out_loc % 16) == 0 && ((uintptr_t)ws % 8) == 0,
                 RONK_EINVAL, "ronk_match_encode: gt_boxes/out_loc must be 16-byte aligned, ws 8-byte aligned");
    EncodeParams p;
    p.cor = (const float4*)h->d_cor;
    p.mcor = (const float4*)h->d_mcor;
    p.enc = (const float4*)h->d_enc;
    p.inside = h->d_inside;
    p.N = h->tab.N;
    p.gt_boxes = (const float4*)gt_boxes;
    p.gt_labels = (const long long*)gt_labels;
    p.gt_counts = gt_counts;
    p.B = B;
    p.Gmax = Gmax;
    p.gcap = (Gmax + 3) & ~3;
    p.high = positive_threshold;
    p.low = ignore_threshold;
    p.ps0 = ps[0]; p.ps1 = ps[1]; p.ps2 = ps[2]; p.ps3 = ps[3];
    p.ignore_between = (match_flags & RONK_MATCH_NO_IGNORE_BETWEEN) ? 0 : 1;
    p.gt_max_first = (match_flags & RONK_MATCH_NO_GT_MAX_FIRST) ? 0 : 1;
    p.out_labels = (long long*)out_labels;
    p.out_loc = (float4*)out_loc;
    p.out_scores = out_scores;
    p.out_matched = out_matched;
    p.out_obj = out_objness;
    p.ws_keys = (u64*)ws;
    p.ws_claimed = (unsigned*)((u64*)ws + (size_t)B * Gmax);

    const int tile = kEncThreads * kEncApt;
    const int tiles = (p.N + tile - 1) / tile;
    // image groups: aim at ~8 resident CTAs per SM, every CTA loops over >= 1 image
    long long want = (long long)h->num_sms * 8;
    int ipc = (int)(((long long)B * tiles) / want);
    if (ipc < 1) ipc = 1;
    int Q = (B + ipc - 1) / ipc;
    if (Q > 65535) Q = 65535;
    size_t smem = (size_t)p.gcap * (16 + 8 + 4);
    cudaStream_t st = (cudaStream_t)stream;
    match_encode_kernel<kEncApt><<<dim3(tiles, Q), kEncThreads, smem, st>>>(p);
    RONK_LAUNCHED();
    match_force_kernel<<<B, 128, (size_t)Gmax * 8, st>>>(p);
    RONK_LAUNCHED();
    return RONK_OK;
}
