// Pieces shared by the two match+encode kernels (grid-structured anchors: match_encode_grid.cu;
// arbitrary flattened anchors: match_encode.cu).  Reference: nets/ssd_common.py:27-147.
#pragma once
#include <math_constants.h>

#include "common.cuh"

namespace ronk {

constexpr int kEncApt = 2;                              // anchors per lane (flat kernel)
constexpr int kEncSet = 32 * kEncApt;                   // anchors per warp-set (flat kernel)

// One work item of the grid kernel: a band of rows of one layer (mode 0: the GT boxes are walked plane by
// plane over the rectangle of cells they intersect) or a flat range with few inside anchors (mode 1: dense
// sweep, lane = inside anchor).  Either way the item writes the outputs of the flat anchors [n_lo, n_hi).
struct GridItem {
    int mode;
    int n_lo, n_hi;
    int r_lo, rows;           // mode 0: first row of the band, rows in the band
    int H, W, A;              // mode 0: layer shape
    int rt_base, ct_base;     // mode 0: first entry of the layer in rowtab / coltab
    int pl_base;              // mode 0: first plane record of the layer
    int layer_n0;             // mode 0: flat index of the layer's first anchor
    int nsub;                 // mode 0: row sub-bands per plane (tasks = A * nsub)
    int pstride;              // mode 0: state entries per plane (rows * W padded)
    int c_lo, c_hi;           // mode 1: compact (inside) anchor range
};

struct EncodeParams {
    const float4* cor;        // [N]   corners of every anchor (force phase)
    const float4* ccor;       // [Nin] corners of the inside anchors
    const int* inside_idx;    // [Nin] flat index of the inside anchors
    const int* cidx;          // [N]   compact index or -1
    const float4* enc;        // [N]   (cy, cx, h', w')
    const uint8_t* inside;    // [N]
    int N, Nin, tiles, anchors_nice;
    const int4* items;        // [tiles] work items {first compact anchor, first flat anchor, end flat anchor, GT split}
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
    // grid kernel only (match_encode_grid.cu)
    const GridItem* gitems;   // [tiles]
    const float4* rowtab;     // per layer [A][H] (ymin, ymax, ymax - ymin, 0) of the second-trip corners
    const float4* coltab;     // per layer [A][W] (xmin, xmax, xmax - xmin, 0)
    const int4* planes;       // per (layer, shape): inside rows [x, y], inside columns [z, w] (x > y: none)
    int key_flat;             // low word of a per-GT key is ~flat anchor (grid kernel) or ~compact anchor (flat kernel)
    u64* ws_keys;             // [B*Gmax] per-GT (iou bits << 32 | ~anchor), zero between calls
    unsigned* ws_claimed;     // [B*Gmax] gt_max_first=False bookkeeping, zero between calls
    unsigned* ws_count;       // [B] tiles finished per image, zero between calls
    const int* order;         // [B] images by descending GT count (grid kernel: heavy images are dispatched first)
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

// branch-free IoU in exactly the reference's op order (ssd_common.py:34-47).
// NICE: every coordinate is 0 or has a magnitude in [2^-15, 2^15] and no GT side exceeds 1 (checked
// per anchor handle and per image), so the quotient takes the inline division sequence and the
// clamps are saturating subtracts; otherwise IEEE div.rn and fmaxf.
template <bool NICE>
__device__ __forceinline__ float iou_ref(float4 t, float ga, float4 a, float aa) {
    float h, w;
    if (NICE) {
        // max(d, 0) as a saturating subtract (one FMA-pipe instruction instead of FADD + FMNMX):
        // exact because d <= the GT side <= 1 (checked per image)
        h = __saturatef(fminf(t.z, a.z) - fmaxf(t.x, a.x));
        w = __saturatef(fminf(t.w, a.w) - fmaxf(t.y, a.y));
    } else {
        h = fmaxf(fminf(t.z, a.z) - fmaxf(t.x, a.x), 0.f);
        w = fmaxf(fminf(t.w, a.w) - fmaxf(t.y, a.y), 0.f);
    }
    float inter = h * w;
    float uni = (ga + aa) - inter;
    // where(union == 0, 0, inter / union); union == 0 implies inter == 0
    return NICE ? div_overlap_nice(inter, uni) : div_overlap(inter, uni);
}

__device__ __forceinline__ bool nice_coord(float v) {
    // 0, or 2^-15 <= |v| <= 2^15 (NaN / inf fail)
    unsigned e = (__float_as_uint(v) >> 23) & 0xffu;
    return v == 0.f || (e >= 127u - 15u && e <= 127u + 15u);
}

// Per image, run by the CTA that finished the image last: g2a[g] = decoded per-GT best anchor
// (all-zero row -> anchor 0); the lowest GT index that claims an anchor wins; score =
// overlap[g, n].  Also restores the workspace to zero for the next call.
__device__ __forceinline__ void force_image(const EncodeParams& p, int b, int G, int* s_n, int* s_cl, const int NT) {
    const int tid = threadIdx.x;
#pragma unroll 1
    for (int g = tid; g < p.Gmax; g += NT) {
        size_t o = (size_t)b * p.Gmax + g;
        u64 key = __ldcg(p.ws_keys + o);
        int n = 0;
        if (key) {
            const unsigned lo = 0xffffffffu - (unsigned)(key & 0xffffffffull);
            n = p.key_flat ? (int)lo : p.inside_idx[lo];
        }
        s_n[g] = n;
        s_cl[g] = p.gt_max_first ? 0 : (int)__ldcg(p.ws_claimed + o);
        if (key) p.ws_keys[o] = 0ull;
        if (!p.gt_max_first) p.ws_claimed[o] = 0u;
    }
    if (tid == 0) p.ws_count[b] = 0u;
    __syncthreads();
#pragma unroll 1
    for (int g0 = 0; g0 < G; g0 += NT) {
        const int g = g0 + tid;
        const bool act = g < G;
        const int n = act ? s_n[g] : -1;
        bool first = act && !s_cl[act ? g : 0];   // gt_max_first=False: a GT that already owns an anchor forces nothing
        const int lim = min(G, g0 + NT);
        for (int g2 = 0; g2 < lim; ++g2) first = first && !(g2 < g && s_n[g2] == n && !s_cl[g2]);
        if (!first) continue;
        const float4 gb = p.gt_boxes[(size_t)b * p.Gmax + g];
        const float4 a = p.cor[n];
        const float4 e = p.enc[n];
        const bool in = p.inside[n] != 0;
        const long long label = p.gt_labels[(size_t)b * p.Gmax + g];
        float iou = iou_ref<false>(gb, (gb.w - gb.y) * (gb.z - gb.x), a, (a.w - a.y) * (a.z - a.x));
        float ov = iou * (in ? 1.f : 0.f);
        size_t o = (size_t)b * p.N + n;
        p.out_labels[o] = label;
        p.out_loc[o] = encode_loc(gb, e, p);
        p.out_scores[o] = ov;
        if (p.out_matched) p.out_matched[o] = g;
        if (p.out_obj) p.out_obj[o] = label > 0 ? 1 : 0;
    }
}

}  // namespace ronk
