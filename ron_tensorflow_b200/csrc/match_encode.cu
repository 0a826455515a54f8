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
constexpr int kEncWarps = kEncThreads / 32;
constexpr int kEncApt = 2;   // anchors per thread

struct EncodeParams {
    const float4* cor;
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
__global__ void __launch_bounds__(kEncThreads)
match_encode_kernel(const __grid_constant__ EncodeParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    float4* s_box = reinterpret_cast<float4*>(smem);
    u64* s_best = reinterpret_cast<u64*>(s_box + p.gcap);
    float* s_area = reinterpret_cast<float*>(s_best + p.gcap);
    int* s_gid = reinterpret_cast<int*>(s_area + p.gcap);
    __shared__ float s_red[4][kEncWarps];
    __shared__ int s_wcnt[kEncWarps];

    const unsigned full = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int warp_n0 = blockIdx.x * (kEncThreads * APT) + warp * (32 * APT);

    // ---- this thread's anchors: corners + area in registers for the whole image loop.
    // Anchors outside the border mask have overlap 0 with everything (ssd_common.py:118):
    // give them an empty box so they never produce a positive intersection.
    float4 a[APT];
    float area[APT];
    float wy0 = CUDART_INF_F, wx0 = CUDART_INF_F, wy1 = -CUDART_INF_F, wx1 = -CUDART_INF_F;
#pragma unroll
    for (int j = 0; j < APT; ++j) {
        int n = warp_n0 + j * 32 + lane;
        a[j] = make_float4(CUDART_INF_F, CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
        area[j] = 0.f;
        if (n < p.N && p.inside[n]) {
            a[j] = p.cor[n];
            area[j] = (a[j].w - a[j].y) * (a[j].z - a[j].x);
            wy0 = fminf(wy0, a[j].x);
            wx0 = fminf(wx0, a[j].y);
            wy1 = fmaxf(wy1, a[j].z);
            wx1 = fmaxf(wx1, a[j].w);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        wy0 = fminf(wy0, __shfl_xor_sync(full, wy0, o));
        wx0 = fminf(wx0, __shfl_xor_sync(full, wx0, o));
        wy1 = fmaxf(wy1, __shfl_xor_sync(full, wy1, o));
        wx1 = fmaxf(wx1, __shfl_xor_sync(full, wx1, o));
    }
    if (lane == 0) {
        s_red[0][warp] = wy0;
        s_red[1][warp] = wx0;
        s_red[2][warp] = wy1;
        s_red[3][warp] = wx1;
    }
    __syncthreads();
    float ty0 = s_red[0][0], tx0 = s_red[1][0], ty1 = s_red[2][0], tx1 = s_red[3][0];
#pragma unroll
    for (int w = 1; w < kEncWarps; ++w) {
        ty0 = fminf(ty0, s_red[0][w]);
        tx0 = fminf(tx0, s_red[1][w]);
        ty1 = fmaxf(ty1, s_red[2][w]);
        tx1 = fmaxf(tx1, s_red[3][w]);
    }

    for (int b = blockIdx.y; b < p.B; b += gridDim.y) {
        int G = p.gt_counts[b];
        G = G < 0 ? 0 : (G > p.Gmax ? p.Gmax : G);
        const float4* gtb = p.gt_boxes + (size_t)b * p.Gmax;

        // ---- order-preserving cull of the GT list against the tile's bounding box.  A pair
        // has a positive intersection only if the GT overlaps the union extent of the tile
        // (float subtraction is sign exact), every other pair contributes exactly 0.
        __syncthreads();
        int ncull = 0;
        for (int g0 = 0; g0 < G; g0 += kEncThreads) {
            int g = g0 + tid;
            bool pass = false;
            float4 gb = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g < G) {
                gb = gtb[g];
                pass = (fminf(gb.z, ty1) > fmaxf(gb.x, ty0)) && (fminf(gb.w, tx1) > fmaxf(gb.y, tx0));
            }
            unsigned bal = __ballot_sync(full, pass);
            if (lane == 0) s_wcnt[warp] = __popc(bal);
            __syncthreads();
            int off = ncull, tot = 0;
#pragma unroll
            for (int w = 0; w < kEncWarps; ++w) {
                int c = s_wcnt[w];
                if (w < warp) off += c;
                tot += c;
            }
            if (pass) {
                int k = off + __popc(bal & ((1u << lane) - 1u));
                s_box[k] = gb;
                s_area[k] = (gb.w - gb.y) * (gb.z - gb.x);
                s_gid[k] = g;
                s_best[k] = 0ull;
            }
            ncull += tot;
            __syncthreads();
        }

        // ---- IoU sweep.  best/bestk: per-anchor running max and FIRST argmax over GT
        // (strict '>' on an ascending GT list == tf.argmax first occurrence).
        float best[APT];
        int bestk[APT];
#pragma unroll
        for (int j = 0; j < APT; ++j) { best[j] = 0.f; bestk[j] = -1; }

        for (int k = 0; k < ncull; ++k) {
            const float4 g = s_box[k];
            // warp-uniform cull against this warp's extent
            if (!((fminf(g.z, wy1) > fmaxf(g.x, wy0)) && (fminf(g.w, wx1) > fmaxf(g.y, wx0)))) continue;
            const float ga = s_area[k];
            unsigned mybits = 0u;
            unsigned myj = 0u;
#pragma unroll
            for (int j = 0; j < APT; ++j) {
                float h = fminf(g.z, a[j].z) - fmaxf(g.x, a[j].x);
                float w = fminf(g.w, a[j].w) - fmaxf(g.y, a[j].y);
                if (h > 0.f && w > 0.f) {
                    float inter = h * w;
                    float uni = (ga + area[j]) - inter;
                    float iou = (uni == 0.f) ? 0.f : inter / uni;
                    if (iou > best[j]) { best[j] = iou; bestk[j] = k; }
                    unsigned bits = __float_as_uint(iou);
                    if (bits > mybits) { mybits = bits; myj = (unsigned)j; }
                }
            }
            // per-GT (max, lowest anchor index): IoU >= 0 so float bits order as integers
            unsigned m = __reduce_max_sync(full, mybits);
            if (m != 0u) {
                unsigned cand = (mybits == m) ? (myj * 32u + (unsigned)lane) : 0xffffffffu;
                unsigned first = __reduce_min_sync(full, cand);
                if (lane == 0) {
                    u64 key = ((u64)m << 32) | (u64)(0xffffffffu - ((unsigned)warp_n0 + first));
                    if (key > s_best[k]) atomicMax(&s_best[k], key);
                }
            }
        }
        __syncthreads();

        for (int k = tid; k < ncull; k += kEncThreads) {
            u64 v = s_best[k];
            if (v != 0ull) atomicMax(p.ws_keys + (size_t)b * p.Gmax + s_gid[k], v);
        }

        // ---- label + encode + store (forced anchors are rewritten by match_force_kernel)
#pragma unroll
        for (int j = 0; j < APT; ++j) {
            int n = warp_n0 + j * 32 + lane;
            if (n >= p.N) continue;
            float mv = best[j];
            int kb = bestk[j];
            int a2g = (kb >= 0) ? s_gid[kb] : 0;
            bool less = mv < p.low;
            bool between = (mv < p.high) && (mv >= p.low);
            bool neg = p.ignore_between ? less : between;
            bool ign = p.ignore_between ? between : less;
            int mi = ign ? -2 : (neg ? -1 : a2g);
            if (G == 0) mi = -1;
            long long label = 0;
            float4 loc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (mi >= 0) {
                float4 gb = (kb >= 0) ? s_box[kb] : gtb[0];
                label = p.gt_labels[(size_t)b * p.Gmax + a2g];
                loc = encode_loc(gb, p.enc[n], p);
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
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        if (s_cl[g]) continue;                    // gt_max_first=False: GT already has an anchor
        const int n = s_n[g];
        bool first = true;
        for (int g2 = 0; g2 < g; ++g2)
            if (s_n[g2] == n && !s_cl[g2]) { first = false; break; }
        if (!first) continue;
        const float4 gb = p.gt_boxes[(size_t)b * p.Gmax + g];
        const float4 a = p.cor[n];
        float h = fmaxf(fminf(gb.z, a.z) - fmaxf(gb.x, a.x), 0.f);
        float w = fmaxf(fminf(gb.w, a.w) - fmaxf(gb.y, a.y), 0.f);
        float inter = h * w;
        float uni = ((gb.w - gb.y) * (gb.z - gb.x) + (a.w - a.y) * (a.z - a.x)) - inter;
        float iou = (uni == 0.f) ? 0.f : inter / uni;
        float ov = iou * (p.inside[n] ? 1.f : 0.f);
        long long label = p.gt_labels[(size_t)b * p.Gmax + g];
        size_t o = (size_t)b * p.N + n;
        p.out_labels[o] = label;
        p.out_loc[o] = encode_loc(gb, p.enc[n], p);
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
    size_t smem = (size_t)p.gcap * (16 + 8 + 4 + 4);
    cudaStream_t st = (cudaStream_t)stream;
    match_encode_kernel<kEncApt><<<dim3(tiles, Q), kEncThreads, smem, st>>>(p);
    RONK_LAUNCHED();
    match_force_kernel<<<B, 128, (size_t)Gmax * 8, st>>>(p);
    RONK_LAUNCHED();
    return RONK_OK;
}
