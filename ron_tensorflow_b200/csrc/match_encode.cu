// K1: fused all-layer GT->anchor matching + target encoding, ONE launch per batch.
//
// Reference: nets/ssd_common.py:27-47 (iou_matrix), :49-75 (do_dual_max_match), :77-147
// (tf_ssd_bboxes_encode_layer), nets/ron_vgg_320.py:686,708 (objectness label).
// Spec: SURVEY.md Appendix A.3/A.4.  Results are bit-exact against oracle/ron_oracle.py.
//
// Work item = (tile of consecutive INSIDE anchors, image).  Anchors outside the border mask
// have overlap exactly 0 with everything (ssd_common.py:118), so only the compacted inside
// anchors (65 % for RON-320) ever enter the IoU sweep; a tile still owns the contiguous range of
// flat anchor indices around its inside anchors and writes all of their outputs coalesced.
// A CTA has 4 warps = SETS anchor sets (64 anchors each, in registers) x SPLIT contiguous parts
// of the image's GT list: (4,1) for large batches, (1,4) when the batch is too small to fill the
// machine, so that the heaviest item (50 GT boxes that all touch the set) is cut in four.
// Per image the CTA
//   1. stages the GT boxes in shared memory;
//   2. sweeps: every warp culls 32 GT boxes at a time against its set's bounding extent (ballot;
//      order preserving, so "first GT wins" needs no extra compare; a culled pair has
//      intersection <= 0, i.e. IoU exactly 0), evaluates IoU in the reference's exact op order
//      two GT boxes per iteration, keeps the per-anchor (max, first argmax) in registers and the
//      per-GT (max, lowest anchor) as a packed u64 in shared memory -- the REDUX + atomicMax
//      path only runs when a lane can reach the current per-GT best;
//   3. labels and stores labels / scores for its flat anchor range; anchors matched by threshold are
//      listed in shared memory and their localisations encoded afterwards, one listed anchor per thread;
//   4. publishes its per-GT bests with one global atomicMax each and bumps the image's tile
//      counter; the CTA that finishes an image LAST applies "lowest GT index claims the anchor"
//      (tf.argmax of the one-hot mask, ssd_common.py:74-75), rewrites those (<= G) anchors and
//      restores the workspace to zero.  No [G,N] matrix ever exists, no second launch.
// Tiles of the coarse layers (big anchors: every GT touches them) are the heaviest and come first
// in the flat order; the grid is tile-major so they are dispatched first.
#include <math_constants.h>
#include <stdlib.h>

#include "encode_common.cuh"

namespace ronk {

// IoU sweep of one warp over GT boxes [g_lo, g_hi).  best/bestg: per-anchor running max and
// FIRST argmax over GT (strict '>' over ascending g == tf.argmax first occurrence).
template <bool NICE>
__device__ __forceinline__ void sweep(const float4* s_box, const float* s_area, u64* s_best, int g_lo, int g_hi,
                                      const float4 (&a)[kEncApt], const float (&area)[kEncApt], float wy0, float wx0,
                                      float wy1, float wx1, unsigned set_c0, float (&best)[kEncApt],
                                      int (&bestg)[kEncApt]) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    for (int g0 = g_lo; g0 < g_hi; g0 += 32) {
        bool touch = false;
        if (g0 + lane < g_hi) {
            float4 t = s_box[g0 + lane];
            touch = (fminf(t.z, wy1) > fmaxf(t.x, wy0)) && (fminf(t.w, wx1) > fmaxf(t.y, wx0));
        }
        unsigned todo = __ballot_sync(full, touch);
        while (todo) {
            // two GT boxes per iteration (the second repeats the first when only one is left:
            // max / first-argmax / atomicMax are idempotent) for instruction-level parallelism
            int gq[2];
            gq[0] = g0 + __ffs(todo) - 1;
            todo &= todo - 1;
            gq[1] = todo ? g0 + __ffs(todo) - 1 : gq[0];
            todo &= todo - 1;
            float4 t[2];
            float ga[2];
            unsigned cur[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                t[q] = s_box[gq[q]];
                ga[q] = s_area[gq[q]];
                // s_best is this warp's private copy (seeded with the image-wide best when the item
                // started); >= 1 so that zero overlaps never pass
                cur[q] = max((unsigned)(s_best[gq[q]] >> 32), 1u);
            }
            unsigned bits[2][kEncApt];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
#pragma unroll
                for (int j = 0; j < kEncApt; ++j)
                    bits[q][j] = __float_as_uint(iou_ref<NICE>(t[q], ga[q], a[j], area[j]));
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                unsigned mybits = 0u;
#pragma unroll
                for (int j = 0; j < kEncApt; ++j) {
                    float iou = __uint_as_float(bits[q][j]);
                    if (iou > best[j]) { best[j] = iou; bestg[j] = gq[q]; }
                    mybits = max(mybits, bits[q][j]);
                }
                // per-GT (max, lowest anchor index): IoU >= 0 so float bits order as integers.
                // Only when some lane reaches the GT's current best does the warp reduce.
                if (__any_sync(full, mybits >= cur[q])) {
                    const unsigned m = __reduce_max_sync(full, mybits);
                    unsigned cand = 0xffffffffu;
#pragma unroll
                    for (int j = kEncApt - 1; j >= 0; --j)
                        if (bits[q][j] == m) cand = (unsigned)(j * 32 + lane);
                    const unsigned first = __reduce_min_sync(full, cand);
                    __syncwarp();                       // every lane has consumed its read of s_best (cur) above
                    if (lane == 0) {
                        u64 key = ((u64)m << 32) | (u64)(0xffffffffu - (set_c0 + first));
                        if (key > s_best[gq[q]]) s_best[gq[q]] = key;       // private to this warp: no atomic
                    }
                    __syncwarp();
                }
            }
        }
    }
}

#ifdef RONK_ENC_TRACE
// profiling build only (tools/enc_trace.py): per CTA 8 words {start ns, SM id, end ns, after staging, after the
// sweep, after the output stores, after the tile counter, unused}
__device__ unsigned long long* g_enc_trace;
__device__ __forceinline__ unsigned long long trace_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define RONK_TRACE_MARK(k)                                                                          \
    do {                                                                                            \
        if (g_enc_trace && threadIdx.x == 0)                                                        \
            g_enc_trace[8 * ((size_t)blockIdx.y * gridDim.x + blockIdx.x) + (k)] = trace_now();     \
    } while (0)
#else
#define RONK_TRACE_MARK(k) do { } while (0)
#endif

constexpr int kEncThreads = 128;
constexpr int kEncPrefetch = 3;     // flat anchors per thread whose compact index is fetched before the sweep

// One work item: SETS sets of 64 inside anchors starting at compact index c0, GT list cut in SPLIT
// parts (SETS * SPLIT = 4 warps); writes the outputs of the flat anchors [n_lo, n_hi).
template <int SETS, int SPLIT>
__device__ __forceinline__ void encode_item(const EncodeParams& p, unsigned char* smem, int b, int c0, int n_lo,
                                            int n_hi) {
    static_assert(SETS * SPLIT * 32 == kEncThreads, "4 warps per CTA");
    constexpr int kTile = SETS * kEncSet;
    float4* s_box = reinterpret_cast<float4*>(smem);              // [gcap] GT corners
    u64* s_best = reinterpret_cast<u64*>(s_box + p.gcap);         // [4][gcap] per warp: per-GT (iou bits, ~compact anchor)
    u64* s_init = s_best + 4 * p.gcap;                            // [gcap] image-wide best when this item started
    long long* s_lab = reinterpret_cast<long long*>(s_init + p.gcap);   // [gcap] GT labels
    float* s_area = reinterpret_cast<float*>(s_lab + p.gcap);     // [gcap]
    __shared__ float s_mv[SPLIT][kTile];                          // per-anchor max overlap, per GT part
    __shared__ int s_mg[SPLIT][kTile];                            // per-anchor first argmax (-1: none)
    __shared__ int s_pn[kTile], s_pg[kTile];                      // anchors matched by threshold: flat index, GT
    __shared__ unsigned s_npos;
    __shared__ int s_last;

    const unsigned full = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int set = warp / SPLIT, part = warp % SPLIT;
    const float4* gtb = p.gt_boxes + (size_t)b * p.Gmax;
    const long long* gtl = p.gt_labels + (size_t)b * p.Gmax;
    const u64* wsk = p.ws_keys + (size_t)b * p.Gmax;
    // Every load whose address does not depend on another load is issued here, up front: the first GT
    // slot of each thread (before the GT count is known), the image-wide bests, this lane's anchors and
    // the compact indices of the first flat anchors it will write.  One global round trip instead of
    // four on the critical path of every CTA.
    if (tid == 0) s_npos = 0u;                                    // ordered before its first use by the barriers below
    const bool spec = tid < p.Gmax;
    const float4 v_spec = spec ? gtb[tid] : make_float4(0.f, 0.f, 0.f, 0.f);
    const long long l_spec = spec ? gtl[tid] : 0ll;
    const u64 k_spec = spec ? __ldcg(wsk + tid) : 0ull;
    int cpre[kEncPrefetch];
#pragma unroll
    for (int q = 0; q < kEncPrefetch; ++q) {
        const int n = n_lo + q * kEncThreads + tid;
        cpre[q] = (n < n_hi) ? p.cidx[n] : -1;
    }
    const int set_c0 = c0 + set * kEncSet;
    float4 a[kEncApt];
    float area[kEncApt];
    float wy0 = CUDART_INF_F, wx0 = CUDART_INF_F, wy1 = -CUDART_INF_F, wx1 = -CUDART_INF_F;
#pragma unroll
    for (int j = 0; j < kEncApt; ++j) {
        int c = set_c0 + j * 32 + lane;
        const bool in = c < p.Nin;
        a[j] = in ? p.ccor[c] : make_float4(CUDART_INF_F, CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
    }
    int G = p.gt_counts[b];
    G = G < 0 ? 0 : (G > p.Gmax ? p.Gmax : G);

    bool nice = true;
    for (int g = tid; g < G; g += kEncThreads) {
        const float4 v = (g == tid) ? v_spec : gtb[g];
        s_box[g] = v;
        s_lab[g] = (g == tid) ? l_spec : gtl[g];
        s_area[g] = (v.w - v.y) * (v.z - v.x);
        // start from the image-wide best published so far (any stale value is a valid lower
        // bound): only overlaps that can still win reach the reduction path
        const u64 k0 = (g == tid) ? k_spec : __ldcg(wsk + g);
#pragma unroll
        for (int w = 0; w < 4; ++w) s_best[w * p.gcap + g] = k0;
        s_init[g] = k0;
        nice = nice && nice_coord(v.x) && nice_coord(v.y) && nice_coord(v.z) && nice_coord(v.w) &&
               (v.z - v.x) <= 1.f && (v.w - v.y) <= 1.f;
    }
#pragma unroll
    for (int j = 0; j < kEncApt; ++j) {
        const bool in = a[j].x != CUDART_INF_F;
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
    nice = __syncthreads_and(nice && p.anchors_nice) != 0;
    RONK_TRACE_MARK(3);

    float best[kEncApt];
    int bestg[kEncApt];
#pragma unroll
    for (int j = 0; j < kEncApt; ++j) { best[j] = 0.f; bestg[j] = -1; }
    const int per = (G + SPLIT - 1) / SPLIT;
    const int g_lo = min(G, part * per), g_hi = min(G, g_lo + per);
    if (nice)
        sweep<true>(s_box, s_area, s_best + warp * p.gcap, g_lo, g_hi, a, area, wy0, wx0, wy1, wx1, (unsigned)set_c0, best, bestg);
    else
        sweep<false>(s_box, s_area, s_best + warp * p.gcap, g_lo, g_hi, a, area, wy0, wx0, wy1, wx1, (unsigned)set_c0, best, bestg);
#pragma unroll
    for (int j = 0; j < kEncApt; ++j) {
        s_mv[part][set * kEncSet + j * 32 + lane] = best[j];
        s_mg[part][set * kEncSet + j * 32 + lane] = bestg[j];
    }
    __syncthreads();
    RONK_TRACE_MARK(4);

    for (int g = tid; g < G; g += kEncThreads) {
        u64 v = s_best[g];
#pragma unroll
        for (int w = 1; w < 4; ++w) v = max(v, s_best[w * p.gcap + g]);
        // only keys this item produced are above the value it started from
        if (v > s_init[g]) atomicMax(p.ws_keys + (size_t)b * p.Gmax + g, v);
    }

    // ---- label + store the item's flat anchor range (forced anchors are rewritten by the CTA that
    // finishes the image last).  Anchors matched by threshold are only listed here: their encoding (~300
    // dependent instructions: 4 IEEE divisions, 2 double-precision logs) runs after the loop, one listed
    // anchor per thread, instead of once per warp iteration that happens to hold one or two of them.
    int q = 0;
    for (int n = n_lo + tid; n < n_hi; n += kEncThreads, ++q) {
        int c;
        switch (q) {                                       // compile-time register selection
            case 0: c = cpre[0]; break;
            case 1: c = cpre[1]; break;
            case 2: c = cpre[2]; break;
            default: c = p.cidx[n]; break;
        }
        float mv = 0.f;
        int a2g = 0;
        if (c >= 0) {
            // parts hold ascending GT ranges: strict '>' keeps the lowest GT index on ties
            int g = -1;
#pragma unroll
            for (int s = 0; s < SPLIT; ++s) {
                float v = s_mv[s][c - c0];
                if (v > mv) { mv = v; g = s_mg[s][c - c0]; }
            }
            a2g = g < 0 ? 0 : g;
        }
        bool less = mv < p.low;
        bool between = (mv < p.high) && (mv >= p.low);
        bool neg = p.ignore_between ? less : between;
        bool ign = p.ignore_between ? between : less;
        int mi = ign ? -2 : (neg ? -1 : a2g);
        if (G == 0) mi = -1;
        long long label = 0;
        if (mi >= 0) {
            label = s_lab[a2g];
            const unsigned slot = atomicAdd(&s_npos, 1u);        // <= kTile: only inside anchors can match
            s_pn[slot] = n;
            s_pg[slot] = a2g;
            if (!p.gt_max_first) atomicOr(p.ws_claimed + (size_t)b * p.Gmax + a2g, 1u);
        } else if (mi < -1) {
            label = -1;
        }
        size_t o = (size_t)b * p.N + n;
        p.out_labels[o] = label;
        if (mi < 0) p.out_loc[o] = make_float4(0.f, 0.f, 0.f, 0.f);
        p.out_scores[o] = mv;
        if (p.out_matched) p.out_matched[o] = mi;
        if (p.out_obj) p.out_obj[o] = label > 0 ? 1 : 0;
    }

    __syncthreads();
    for (unsigned k = tid; k < s_npos; k += kEncThreads) {
        const int n = s_pn[k];
        p.out_loc[(size_t)b * p.N + n] = encode_loc(s_box[s_pg[k]], p.enc[n], p);
    }

    // ---- publish; the last item of the image applies the per-GT forcing.  Barrier first, then
    // one thread fences and bumps the counter (release pattern of a grid-wide barrier).
    __syncthreads();
    RONK_TRACE_MARK(5);
    if (tid == 0) {
        __threadfence();
        unsigned prev = atomicAdd(p.ws_count + b, 1u);
        s_last = (prev == (unsigned)p.tiles - 1u) ? 1 : 0;
    }
    __syncthreads();
    RONK_TRACE_MARK(6);
    if (s_last) {
        __threadfence();
        // s_best (8 B per GT slot) is free now: reuse it as two int arrays
        force_image(p, b, G, reinterpret_cast<int*>(s_best), reinterpret_cast<int*>(s_best) + p.gcap, kEncThreads);
    }
}

// grid (B, items): blockIdx.y walks the handle's work-item table (heavy, coarse-layer items first)
__global__ void __launch_bounds__(kEncThreads, 8)
match_encode_kernel(const __grid_constant__ EncodeParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
#ifdef RONK_ENC_TRACE
    unsigned long long* tr = g_enc_trace ? g_enc_trace + 8 * ((size_t)blockIdx.y * gridDim.x + blockIdx.x) : nullptr;
    if (tr && threadIdx.x == 0) {
        unsigned sm;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        tr[0] = trace_now();
        tr[1] = sm;
    }
#endif
    const int4 item = __ldg(p.items + blockIdx.y);
    if (item.w == 4)
        encode_item<1, 4>(p, smem, blockIdx.x, item.x, item.y, item.z);
    else
        encode_item<4, 1>(p, smem, blockIdx.x, item.x, item.y, item.z);
#ifdef RONK_ENC_TRACE
    __syncthreads();
    if (tr && threadIdx.x == 0) tr[2] = trace_now();
#endif
}

__global__ void zero_u32_kernel(unsigned* p, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0u;
}

}  // namespace ronk

namespace ronk {
int launch_match_encode_grid(const ronk_anchors* h, EncodeParams& p, int B, cudaStream_t st);   // match_encode_grid.cu
}

using namespace ronk;

#ifdef RONK_ENC_TRACE
extern "C" int ronk_debug_set_enc_trace(void* buf) {
    RONK_CUDA(cudaMemcpyToSymbol(g_enc_trace, &buf, sizeof(buf)));
    return RONK_OK;
}
#endif

extern "C" size_t ronk_encode_workspace_bytes(int B, int Gmax) {
    if (B < 1 || Gmax < 1) return 0;
    // per-GT keys + claimed flags, per-image item counters (all zero between calls), image order (scratch)
    return (size_t)B * Gmax * (sizeof(u64) + sizeof(unsigned)) + (size_t)B * sizeof(unsigned) + (size_t)B * sizeof(int);
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
    p.ccor = (const float4*)h->d_ccor;
    p.inside_idx = h->d_inside_idx;
    p.cidx = h->d_cidx;
    p.enc = (const float4*)h->d_enc;
    p.inside = h->d_inside;
    p.N = h->tab.N;
    p.Nin = h->n_inside;
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
    p.ws_count = p.ws_claimed + (size_t)B * Gmax;
    p.order = (const int*)(p.ws_count + B);

    p.anchors_nice = h->anchors_nice;
    p.gitems = nullptr;
    p.rowtab = p.coltab = nullptr;
    p.planes = nullptr;
    p.key_flat = 0;
    // Anchors on the regular grids the reference generates can take the grid kernel (match_encode_grid.cu): fewer
    // instructions per image (exact rectangles of intersecting cells instead of a culled dense sweep), but a CTA owns
    // a whole band of a layer and lives 2-3x longer, so it only pays once the batch fills the machine several times
    // over (measured crossover for RON-320: ~110 images).  Arbitrary flattened anchors (ronk_anchors_create_flat)
    // and small batches take the generic kernel below.
    {
        int use_grid = h->grid_ok && (long long)B * h->tab.N >= 2400000ll;
        if (const char* e = getenv("RONK_ENC_KERNEL")) {           // tuning knob: "grid" / "generic"
            if (e[0] == 'g' && e[1] == 'r') use_grid = h->grid_ok;
            else if (e[0] == 'g' && e[1] == 'e') use_grid = 0;
        }
        if (use_grid) return launch_match_encode_grid(h, p, B, (cudaStream_t)stream);
    }
    // Work-item table: the batch decides how finely the heaviest items are cut.  A tiny batch cannot
    // fill 148 SMs x 40 warps with whole-GT-list items, and the kernel then lasts as long as its
    // slowest CTA (a coarse-layer tile against 50 GT boxes): cut those (table 1) or everything
    // (table 2) into 4 GT parts.  Large batches keep whole items (table 0): least overhead.
    const long long sets = (long long)B * ((p.Nin + kEncSet - 1) / kEncSet), slots = (long long)h->num_sms * 40;
    // measured crossovers for RON-320 (tools/enc_table_sweep.py): batch ~13 and ~55 with up to 50 GT boxes per image; with
    // 100-200 boxes per image the heavy items are 2-4 x longer and cutting everything pays up to batch ~27
    const long long fine = Gmax >= 100 ? slots : slots / 2;
    int table = sets < fine ? 2 : (sets < slots * 2 ? 1 : 0);
    if (const char* e = getenv("RONK_ENC_TABLE")) {            // tuning knob
        int v = atoi(e);
        if (v >= 0 && v <= 2) table = v;
    }
    p.items = (const int4*)h->d_items[table];
    p.tiles = h->n_items[table];
    RONK_REQUIRE(p.tiles <= 65535, RONK_ELIMIT, "ronk_match_encode: too many anchors for one launch");
    size_t smem = (size_t)p.gcap * (16 + 4 * 8 + 8 + 8 + 4);
    cudaStream_t st = (cudaStream_t)stream;
    if (smem > 48 * 1024)
        RONK_CUDA(cudaFuncSetAttribute(match_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    match_encode_kernel<<<dim3((unsigned)B, (unsigned)p.tiles), kEncThreads, smem, st>>>(p);
    RONK_LAUNCHED();
    return RONK_OK;
}
