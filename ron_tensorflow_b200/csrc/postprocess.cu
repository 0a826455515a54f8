// K2: decode + objectness gate + score threshold + clip + min-size + per-class segmented top-k.
//
// Reference: nets/ssd_common.py:448-498 (decode), eval_ron_network.py:227-229 (objectness gate),
// nets/ssd_common.py:504-589 (select), tf_extended/bboxes.py:105-144 (clip),
// nets/ron_vgg_320.py:196-233 (bboxes_filter_min), tf_extended/bboxes.py:60-101 (bboxes_sort).
// Spec: SURVEY.md Appendix A.5/A.6.  Bit-exact against oracle/ron_oracle.py.
//
//   scatter_candidates_kernel  persistent CTAs stream [256 anchors x C] class-score tiles into
//     shared memory with 1-D TMA bulk copies (cp.async.bulk + mbarrier, double buffered), so
//     every score row is read from HBM exactly once for ALL classes; each thread gates its
//     anchor by objectness, thresholds the C-1 class scores, decodes + clips + size-tests the
//     box once if any class survives, and appends u64 keys (score bits << 32 | ~anchor) to
//     per-(image, class) candidate lists with warp-ballot + one global atomic per class per CTA.
//   select_topk_kernel         one CTA per (image, class): exact radix select of the top K keys,
//     bitonic sort, box re-decode of the K winners.
#include "common.cuh"
#include "topk.cuh"

namespace ronk {

constexpr int kTileRows = 256;

struct PostParams {
    LayerTable tab;
    const float* loc[kMaxLayers];
    const float* cls[kMaxLayers];
    const float* obj[kMaxLayers];    // all NULL when has_obj == 0
    int tile_off[kMaxLayers + 1];    // per image: first tile of each layer
    const float4* dec;               // decode anchors (y, x, h, w)
    int B, C, K, has_obj, has_clip, decoded;
    float obj_thr, sel_thr, min_size;
    float4 clip;
    float ps0, ps1, ps2, ps3;
    int* counts;                     // [B*(C-1)]
    u64* keys;                       // [B*(C-1), cap]
    int cap;
    long long num_tiles;
    float* out_scores;
    float4* out_boxes;
    int* out_idx;
};

// ----------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u64* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(u64* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, u64* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct TileInfo {
    int b, layer, t0, rows, n_l;
    long long start_w, a0, a1, end_w;   // word offsets inside the layer's class array
};

__device__ __forceinline__ TileInfo tile_info(const PostParams& p, long long tile) {
    TileInfo t;
    const int tpi = p.tile_off[p.tab.L];
    t.b = (int)(tile / tpi);
    int r = (int)(tile % tpi);
    int l = 0;
    while (l + 1 < p.tab.L && r >= p.tile_off[l + 1]) ++l;
    t.layer = l;
    t.n_l = p.tab.offs[l + 1] - p.tab.offs[l];
    t.t0 = (r - p.tile_off[l]) * kTileRows;
    t.rows = min(kTileRows, t.n_l - t.t0);
    t.start_w = ((long long)t.b * t.n_l + t.t0) * p.C;
    t.end_w = t.start_w + (long long)t.rows * p.C;
    long long total_w = (long long)p.B * t.n_l * p.C;
    t.a0 = t.start_w & ~3ll;
    long long a1 = (t.end_w + 3) & ~3ll;
    long long lim = total_w & ~3ll;
    t.a1 = a1 < lim ? a1 : lim;
    if (t.a1 < t.a0) t.a1 = t.a0;
    return t;
}

__global__ void __launch_bounds__(kTileRows)
scatter_candidates_kernel(const __grid_constant__ PostParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int C = p.C, CM = p.C - 1;
    const int stage_floats = (kTileRows * C + 4 + 3) & ~3;
    float* const s_stage0 = reinterpret_cast<float*>(smem);
    int* s_wcnt = reinterpret_cast<int*>(reinterpret_cast<float*>(smem) + 2 * stage_floats);   // [8][CM]
    int* s_wbase = s_wcnt + 8 * CM;                                                           // [8][CM]
    __shared__ __align__(8) u64 s_bar[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        fence_proxy_async();
    }
    __syncthreads();

    long long tile = blockIdx.x;
    if (tile >= p.num_tiles) return;
    if (tid == 0) {
        TileInfo t = tile_info(p, tile);
        unsigned bytes = (unsigned)((t.a1 - t.a0) * 4);
        if (bytes) {
            mbar_arrive_expect_tx(&s_bar[0], bytes);
            tma_load_1d(s_stage0, p.cls[t.layer] + t.a0, bytes, &s_bar[0]);
        }
    }
    unsigned phase = 0u;   // bit st = parity the next wait on stage st must see
    for (int it = 0; tile < p.num_tiles; ++it, tile += gridDim.x) {
        const int st = it & 1;
        // prefetch the next tile into the other stage (freed by the barrier that ended iteration it-1)
        const long long next = tile + gridDim.x;
        if (tid == 0 && next < p.num_tiles) {
            TileInfo tn = tile_info(p, next);
            unsigned bytes = (unsigned)((tn.a1 - tn.a0) * 4);
            if (bytes) {
                mbar_arrive_expect_tx(&s_bar[st ^ 1], bytes);
                tma_load_1d(s_stage0 + (st ^ 1) * stage_floats, p.cls[tn.layer] + tn.a0, bytes, &s_bar[st ^ 1]);
            }
        }
        const TileInfo t = tile_info(p, tile);
        const int nl = t.t0 + tid;
        const bool rowok = tid < t.rows;
        const int n = p.tab.offs[t.layer] + nl;
        const size_t row = (size_t)t.b * t.n_l + nl;

        // independent global loads first, so they overlap the wait for the bulk copy
        bool gate = rowok;
        float4 l4 = make_float4(0.f, 0.f, 0.f, 0.f), a4 = l4;
        if (rowok) {
            if (p.has_obj) gate = p.obj[t.layer][row] > p.obj_thr;
            if (gate) {
                l4 = reinterpret_cast<const float4*>(p.loc[t.layer])[row];
                a4 = p.dec[n];
            }
        }
        if (t.a1 > t.a0) {
            mbar_wait(&s_bar[st], (phase >> st) & 1u);
            phase ^= 1u << st;
        }

        const float* srow = s_stage0 + st * stage_floats + (t.start_w - t.a0) + (long long)tid * C;
        const float* grow = p.cls[t.layer] + t.start_w + (long long)tid * C;
        const bool tail = t.end_w > t.a1;   // last <=3 words of the layer array are not covered by the bulk copy
        const long long wlim = t.a1 - t.start_w - (long long)tid * C;   // words of this row that sit in smem

        bool any = false;
        if (gate) {
            for (int c = 1; c < C; ++c) {
                float v = (!tail || c < wlim) ? srow[c] : grow[c];
                any |= v > p.sel_thr;
            }
        }
        bool valid = false;
        u64 nkey = (u64)(0xffffffffu - (unsigned)n);
        if (any) {
            float4 box = p.decoded ? l4 : decode_box(l4, a4, p.ps0, p.ps1, p.ps2, p.ps3);
            if (p.has_clip) box = clip_box(box, p.clip);
            valid = true;
            if (p.min_size >= 0.f) {
                float h = box.z - box.x;
                float w = box.w - box.y;
                valid = (w > p.min_size) && (h > p.min_size);
            }
        }
        // phase 1: per-warp, per-class candidate counts
        for (int c = 1; c < C; ++c) {
            bool pass = false;
            if (valid) {
                float v = (!tail || c < wlim) ? srow[c] : grow[c];
                pass = v > p.sel_thr;
            }
            unsigned bal = __ballot_sync(0xffffffffu, pass);
            if (lane == 0) s_wcnt[warp * CM + (c - 1)] = __popc(bal);
        }
        __syncthreads();
        // one global atomic per class per CTA reserves the tile's slots in the (image, class) list
        for (int c = tid; c < CM; c += kTileRows) {
            int tot = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) tot += s_wcnt[w * CM + c];
            int base = 0;
            if (tot) base = atomicAdd(p.counts + (size_t)t.b * CM + c, tot);
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                s_wbase[w * CM + c] = base;
                base += s_wcnt[w * CM + c];
            }
        }
        __syncthreads();
        // phase 2: write keys
        for (int c = 1; c < C; ++c) {
            bool pass = false;
            float v = 0.f;
            if (valid) {
                v = (!tail || c < wlim) ? srow[c] : grow[c];
                pass = v > p.sel_thr;
            }
            unsigned bal = __ballot_sync(0xffffffffu, pass);
            if (pass) {
                int pos = s_wbase[warp * CM + (c - 1)] + __popc(bal & ((1u << lane) - 1u));
                if (pos < p.cap)
                    p.keys[((size_t)t.b * CM + (c - 1)) * p.cap + pos] = ((u64)__float_as_uint(v) << 32) | nkey;
            }
        }
        __syncthreads();   // stage st and the count arrays are free again
    }
}

// key source: staged prefix in shared memory, remainder straight from the global list
struct ListSrc {
    const u64* g;
    const u64* s;
    int staged;
    __device__ __forceinline__ u64 get(int i) const { return i < staged ? s[i] : g[i]; }
};

__device__ __forceinline__ float4 redecode_box(const PostParams& p, int b, int n) {
    int l = layer_of(p.tab, n);
    int n_l = p.tab.offs[l + 1] - p.tab.offs[l];
    size_t row = (size_t)b * n_l + (n - p.tab.offs[l]);
    float4 l4 = reinterpret_cast<const float4*>(p.loc[l])[row];
    float4 box = p.decoded ? l4 : decode_box(l4, p.dec[n], p.ps0, p.ps1, p.ps2, p.ps3);
    if (p.has_clip) box = clip_box(box, p.clip);
    return box;
}

constexpr int kStageKeys = 4096;

__global__ void __launch_bounds__(kTopkThreads)
select_topk_kernel(const __grid_constant__ PostParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    u64* s_sort = reinterpret_cast<u64*>(smem);                 // pow2(K)
    u64* s_keys = s_sort + next_pow2(p.K);                      // kStageKeys
    __shared__ unsigned s_hist[256];
    __shared__ int s_ctl[4];
    const int seg = blockIdx.x;
    const int CM = p.C - 1;
    const int b = seg / CM;
    int n = p.counts[seg];
    n = n > p.cap ? p.cap : n;
    const u64* g = p.keys + (size_t)seg * p.cap;
    const int staged = n < kStageKeys ? n : kStageKeys;
    for (int i = threadIdx.x; i < staged; i += kTopkThreads) s_keys[i] = g[i];
    __syncthreads();
    ListSrc src{g, s_keys, staged};
    block_topk_sorted(src, n, p.K, s_hist, s_ctl, s_sort);
    for (int r = threadIdx.x; r < p.K; r += kTopkThreads) {
        u64 k = s_sort[r];
        float sc = 0.f;
        float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
        int idx = -1;
        if (k != 0ull) {
            sc = __uint_as_float((unsigned)(k >> 32));
            idx = (int)(0xffffffffu - (unsigned)(k & 0xffffffffull));
            box = redecode_box(p, b, idx);
        }
        size_t o = (size_t)seg * p.K + r;
        p.out_scores[o] = sc;
        p.out_boxes[o] = box;
        if (p.out_idx) p.out_idx[o] = idx;
    }
}

__global__ void zero_i32_kernel(int* p, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0;
}

// ------------------------------------------------------------------ stand-alone decode / clip
__global__ void __launch_bounds__(256)
decode_kernel(const float4* __restrict__ loc, const float4* __restrict__ dec, int first, int n, long long total,
              float ps0, float ps1, float ps2, float ps3, float4* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int a = first + (int)(i % n);
    out[i] = decode_box(loc[i], dec[a], ps0, ps1, ps2, ps3);
}

__global__ void __launch_bounds__(256)
clip_kernel(const float4* __restrict__ in, long long n, float4 ref, float4* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = clip_box(in[i], ref);
}

// ------------------------------------------------------------------ generic bboxes_sort
__device__ __forceinline__ unsigned orderable(float s) {
    unsigned u = __float_as_uint(s);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unorderable(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

struct ScoreSrc {
    const float* g;
    __device__ __forceinline__ u64 get(int i) const {
        return ((u64)orderable(g[i]) << 32) | (u64)(0xffffffffu - (unsigned)i);
    }
};

__global__ void __launch_bounds__(kTopkThreads)
sort_topk_kernel(const float* __restrict__ scores, const float4* __restrict__ boxes, int N, int K,
                 float* __restrict__ out_scores, float4* __restrict__ out_boxes, int* __restrict__ out_idx) {
    extern __shared__ __align__(128) unsigned char smem[];
    u64* s_sort = reinterpret_cast<u64*>(smem);
    __shared__ unsigned s_hist[256];
    __shared__ int s_ctl[4];
    const size_t row = blockIdx.x;
    ScoreSrc src{scores + row * N};
    block_topk_sorted(src, N, K, s_hist, s_ctl, s_sort);
    for (int r = threadIdx.x; r < K; r += kTopkThreads) {
        u64 k = s_sort[r];
        float sc = 0.f;
        float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
        int idx = -1;
        if (k != 0ull) {
            sc = unorderable((unsigned)(k >> 32));
            idx = (int)(0xffffffffu - (unsigned)(k & 0xffffffffull));
            box = boxes[row * N + idx];
        }
        out_scores[row * K + r] = sc;
        out_boxes[row * K + r] = box;
        if (out_idx) out_idx[row * K + r] = idx;
    }
}

}  // namespace ronk

using namespace ronk;

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" size_t ronk_select_workspace_bytes(const ronk_anchors_t* h, int B, int C, int K) {
    if (!h || B < 1 || C < 2 || K < 1) return 0;
    size_t segs = (size_t)B * (C - 1);
    return align_up(segs * sizeof(int), 256) + segs * (size_t)h->tab.N * sizeof(u64);
}

extern "C" int ronk_decode_select_topk(const ronk_anchors_t* h, const float* const* loc_layers,
                                       const float* const* cls_layers, const float* const* obj_layers, int B,
                                       int C, float objectness_threshold, float select_threshold,
                                       const float* clip, float min_size, const float* ps, int K,
                                       int select_flags, float* out_scores, float* out_boxes, int32_t* out_idx, void* ws,
                                       void* stream) {
    RONK_REQUIRE(h != nullptr, RONK_EINVAL, "ronk_decode_select_topk: NULL anchor handle");
    RONK_REQUIRE(loc_layers && cls_layers && ps && out_scores && out_boxes && ws, RONK_EINVAL,
                 "ronk_decode_select_topk: NULL pointer argument");
    RONK_REQUIRE(B >= 1 && C >= 2 && C <= 1024, RONK_EINVAL, "ronk_decode_select_topk: need B >= 1 and 2 <= C <= 1024");
    RONK_REQUIRE(K >= 1 && K <= 16384, RONK_ELIMIT, "ronk_decode_select_topk: need 1 <= K <= 16384");
    RONK_REQUIRE(select_threshold >= 0.f, RONK_EINVAL,
                 "ronk_decode_select_topk: select_threshold must be >= 0 (None in the reference is 0)");
    RONK_REQUIRE(((uintptr_t)out_boxes % 16) == 0 && ((uintptr_t)ws % 256) == 0, RONK_EINVAL,
                 "ronk_decode_select_topk: out_boxes must be 16-byte and ws 256-byte aligned");
    PostParams p;
    p.tab = h->tab;
    int toff = 0;
    for (int l = 0; l < h->tab.L; ++l) {
        RONK_REQUIRE(loc_layers[l] && cls_layers[l] && (!obj_layers || obj_layers[l]), RONK_EINVAL,
                     "ronk_decode_select_topk: NULL layer pointer");
        RONK_REQUIRE(((uintptr_t)loc_layers[l] % 16) == 0 && ((uintptr_t)cls_layers[l] % 16) == 0, RONK_EINVAL,
                     "ronk_decode_select_topk: layer pointers must be 16-byte aligned");
        p.loc[l] = loc_layers[l];
        p.cls[l] = cls_layers[l];
        p.obj[l] = obj_layers ? obj_layers[l] : nullptr;
        p.tile_off[l] = toff;
        int n_l = h->tab.offs[l + 1] - h->tab.offs[l];
        toff += (n_l + kTileRows - 1) / kTileRows;
    }
    p.tile_off[h->tab.L] = toff;
    p.dec = (const float4*)h->d_dec;
    p.B = B;
    p.C = C;
    p.K = K;
    p.has_obj = obj_layers ? 1 : 0;
    p.has_clip = clip ? 1 : 0;
    p.decoded = (select_flags & RONK_SELECT_LOC_DECODED) ? 1 : 0;
    p.obj_thr = objectness_threshold;
    p.sel_thr = select_threshold;
    p.min_size = min_size;
    p.clip = clip ? make_float4(clip[0], clip[1], clip[2], clip[3]) : make_float4(0.f, 0.f, 0.f, 0.f);
    p.ps0 = ps[0]; p.ps1 = ps[1]; p.ps2 = ps[2]; p.ps3 = ps[3];
    const size_t segs = (size_t)B * (C - 1);
    p.counts = (int*)ws;
    p.keys = (u64*)((char*)ws + align_up(segs * sizeof(int), 256));
    p.cap = h->tab.N;
    p.num_tiles = (long long)toff * B;
    p.out_scores = out_scores;
    p.out_boxes = (float4*)out_boxes;
    p.out_idx = out_idx;
    cudaStream_t st = (cudaStream_t)stream;

    zero_i32_kernel<<<(unsigned)((segs + 255) / 256), 256, 0, st>>>(p.counts, segs);
    RONK_LAUNCHED();

    const int stage_floats = (kTileRows * C + 4 + 3) & ~3;
    size_t smem_a = (size_t)2 * stage_floats * 4 + (size_t)16 * (C - 1) * 4;
    RONK_REQUIRE(smem_a <= 220 * 1024, RONK_ELIMIT, "ronk_decode_select_topk: C too large for the shared-memory tile");
    if (smem_a > 48 * 1024)
        RONK_CUDA(cudaFuncSetAttribute(scatter_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
    int per_sm = (int)((220 * 1024) / (smem_a + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    long long grid = (long long)h->num_sms * per_sm;
    if (grid > p.num_tiles) grid = p.num_tiles;
    scatter_candidates_kernel<<<(unsigned)grid, kTileRows, smem_a, st>>>(p);
    RONK_LAUNCHED();

    int P = 1;
    while (P < K) P <<= 1;
    size_t smem_b = (size_t)(P + kStageKeys) * sizeof(u64);
    if (smem_b > 48 * 1024)
        RONK_CUDA(cudaFuncSetAttribute(select_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
    select_topk_kernel<<<(unsigned)segs, kTopkThreads, smem_b, st>>>(p);
    RONK_LAUNCHED();
    return RONK_OK;
}

extern "C" int ronk_decode(const ronk_anchors_t* h, const float* loc, int B, int first_anchor, int n,
                           const float* ps, float* out_boxes, void* stream) {
    RONK_REQUIRE(h && loc && ps && out_boxes, RONK_EINVAL, "ronk_decode: NULL argument");
    RONK_REQUIRE(B >= 1 && n >= 1 && first_anchor >= 0 && first_anchor + n <= h->tab.N, RONK_EINVAL,
                 "ronk_decode: anchor range outside the handle");
    RONK_REQUIRE(((uintptr_t)loc % 16) == 0 && ((uintptr_t)out_boxes % 16) == 0, RONK_EINVAL,
                 "ronk_decode: pointers must be 16-byte aligned");
    long long total = (long long)B * n;
    decode_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)loc, (const float4*)h->d_dec, first_anchor, n, total, ps[0], ps[1], ps[2], ps[3],
        (float4*)out_boxes);
    RONK_LAUNCHED();
    return RONK_OK;
}

extern "C" int ronk_clip(const float* clip, const float* boxes, long long n, float* out_boxes, void* stream) {
    RONK_REQUIRE(clip && boxes && out_boxes && n >= 0, RONK_EINVAL, "ronk_clip: bad argument");
    RONK_REQUIRE(((uintptr_t)boxes % 16) == 0 && ((uintptr_t)out_boxes % 16) == 0, RONK_EINVAL,
                 "ronk_clip: pointers must be 16-byte aligned");
    if (n == 0) return RONK_OK;
    clip_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)boxes, n, make_float4(clip[0], clip[1], clip[2], clip[3]), (float4*)out_boxes);
    RONK_LAUNCHED();
    return RONK_OK;
}

extern "C" int ronk_sort_topk(const float* scores, const float* boxes, int S, int N, int K, float* out_scores,
                              float* out_boxes, int32_t* out_idx, void* stream) {
    RONK_REQUIRE(scores && boxes && out_scores && out_boxes, RONK_EINVAL, "ronk_sort_topk: NULL argument");
    RONK_REQUIRE(S >= 1 && N >= 0 && K >= 1 && K <= 16384, RONK_ELIMIT, "ronk_sort_topk: need S >= 1, 1 <= K <= 16384");
    RONK_REQUIRE(((uintptr_t)boxes % 16) == 0 && ((uintptr_t)out_boxes % 16) == 0, RONK_EINVAL,
                 "ronk_sort_topk: box pointers must be 16-byte aligned");
    int P = 1;
    while (P < K) P <<= 1;
    size_t smem = (size_t)P * sizeof(u64);
    if (smem > 48 * 1024)
        RONK_CUDA(cudaFuncSetAttribute(sort_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sort_topk_kernel<<<S, kTopkThreads, smem, (cudaStream_t)stream>>>(scores, (const float4*)boxes, N, K, out_scores,
                                                                     (float4*)out_boxes, out_idx);
    RONK_LAUNCHED();
    return RONK_OK;
}
