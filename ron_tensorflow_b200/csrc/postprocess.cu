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
#include <math.h>

#include <stdlib.h>
#include "common.cuh"
#include "topk.cuh"

#ifndef RONK_TOPK_MINB
#define RONK_TOPK_MINB 6   // 40 registers; 8 (32 registers) spills 104 B
#endif

namespace ronk {


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
    float* thr;                      // [B*(C-1)] score threshold of every (image, class): sel_thr, or a sampled pivot
    u64* keys;                       // [B*(C-1), cap]
    float4* boxes;                   // [B, N] final (decoded, clipped) box of every surviving anchor
    int cap;
    long long num_tiles;
    // tile subset of this scatter launch: permuted tile index r' in [r_lo, r_hi) of every image,
    // (the handle's tile table is stored in that order: r' -> tile (r' * m) % tiles, m ~ 0.618 tiles, coprime)
    int r_lo, r_hi;
    const int4* tile_tab;            // [tiles per image] permuted tile table of the anchor handle
    int force_rebuild;               // test hook: always take the exact-rebuild path of pivoted segments
    int pivot_rank;                  // rank (from the top) of the sampled key that becomes the pivot; 0 = no sampling
    float* out_scores;
    float4* out_boxes;
    int* out_idx;
    const int* only_flagged;         // select_topk_kernel: when given, segments whose flag is 0 are left alone
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

// tile rp (permuted index in [0, tiles per image)) of image b
__device__ __forceinline__ TileInfo tile_info(const PostParams& p, int b, int rp) {
    TileInfo t;
    const int4 e = __ldg(p.tile_tab + rp);
    t.b = b;
    t.layer = e.x;
    t.t0 = e.y;
    t.rows = e.z;
    t.n_l = e.w;
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

constexpr int kScatWarps = kTileRows / 32;

// Class scan of one tile: lane = class, warp w walks every 8th surviving valid row.  The class
// threshold sits in a register, a row's scores are consecutive words of shared memory (no bank
// conflict), and nothing has to be ordered -- keys of one (image, class) list may land in any order,
// they are sorted later -- so there are no ballots or prefix sums: pass 1 counts per lane, ONE
// atomic per lane reserves the slots of all (up to 32) classes at once, pass 2 stores the u64 keys
// (score bits << 32 | ~anchor).  TAIL: the last <= 3 score words of a layer array are not covered by
// the 16-byte granular bulk copy and are read from global memory.
// Fast form (every score word of the tile is in shared memory).  s_voff[j] = word offset of the j-th
// surviving row inside the tile, or the offset of a row of -inf for a row whose box failed the size
// test, so pass 1 is branch-free: one load of the (warp-uniform) offset, one of the score, a compare,
// and a shift-in of the result bit.  Bit (nq - 1 - q) of `hit` belongs to the warp's q-th row.
__device__ __forceinline__ void class_scan_fast(const PostParams& p, const TileInfo& t, const float* s_cls,
                                                const int* s_voff, const int* s_vrow, int n_rows) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int CM = p.C - 1;
    const unsigned nkey0 = 0xffffffffu - (unsigned)(p.tab.offs[t.layer] + t.t0);
    const size_t seg0 = (size_t)t.b * CM;
    const int nq = n_rows > warp ? (n_rows - warp + kScatWarps - 1) / kScatWarps : 0;
    for (int cb = 0; cb < CM; cb += 32) {
        const bool act = cb + lane < CM;
        const int ci = act ? cb + lane : 0;
        const float thr = act ? p.thr[seg0 + ci] : __int_as_float(0x7f800000);
        const float* col = s_cls + 1 + ci;
        unsigned hit = 0u;
#pragma unroll 4
        for (int j = warp; j < n_rows; j += kScatWarps) {
            const float v = col[s_voff[j]];
            hit = hit + hit + (v > thr ? 1u : 0u);
        }
        if (hit == 0u) continue;
        u64* dst = p.keys + (seg0 + ci) * p.cap + atomicAdd(p.counts + seg0 + ci, __popc(hit));
        while (hit) {
            const int bp = __ffs(hit) - 1;
            hit &= hit - 1;
            const int j = warp + (nq - 1 - bp) * kScatWarps;
            const float v = col[s_voff[j]];
            *dst++ = ((u64)__float_as_uint(v) << 32) | (u64)(nkey0 - (unsigned)s_vrow[j]);
        }
    }
}

template <bool TAIL>
__device__ __forceinline__ void class_scan(const PostParams& p, const TileInfo& t, const float* s_cls,
                                           const int* s_vrows, int nslots) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C = p.C, CM = p.C - 1;
    const unsigned nkey0 = 0xffffffffu - (unsigned)(p.tab.offs[t.layer] + t.t0);
    const int wlim = (int)(t.a1 - t.start_w);   // words of this tile that sit in smem
    const float* gcls = p.cls[t.layer] + t.start_w;
    const size_t seg0 = (size_t)t.b * CM;
    for (int cb = 0; cb < CM; cb += 32) {
        const bool act = cb + lane < CM;
        const int ci = act ? cb + lane : 0;
        const float thr = act ? p.thr[seg0 + ci] : __int_as_float(0x7f800000);
        const float* col = s_cls + 1 + ci;
        // pass 1: bit q of `hit` = this lane's class passes on the warp's q-th row (<= 32 rows per warp)
        unsigned hit = 0u;
        {
            int q = 0;
            for (int j = warp; j < nslots; j += kScatWarps, ++q) {
                const int r = s_vrows[j];                  // warp-uniform; -1 = empty slot
                if (r < 0) continue;
                const int e = r * C;
                const float v = (!TAIL || e + 1 + ci < wlim) ? col[e] : gcls[e + 1 + ci];
                if (v > thr) hit |= 1u << q;
            }
        }
        if (hit == 0u) continue;
        u64* dst = p.keys + (seg0 + ci) * p.cap + atomicAdd(p.counts + seg0 + ci, __popc(hit));
        // pass 2: only the rows that passed
        while (hit) {
            const int q = __ffs(hit) - 1;
            hit &= hit - 1;
            const int r = s_vrows[warp + q * kScatWarps];
            const int e = r * C;
            const float v = (!TAIL || e + 1 + ci < wlim) ? col[e] : gcls[e + 1 + ci];
            *dst++ = ((u64)__float_as_uint(v) << 32) | (u64)(nkey0 - (unsigned)r);
        }
    }
}

// One tile = 256 consecutive anchors of one layer of one image, all C class scores.
//  A. every thread gates its own row by objectness; the surviving rows are compacted
//     (ballot + warp prefix) into s_rows, so that the two expensive steps below run on full warps;
//  B. one thread per SURVIVING row decodes + clips + size-tests the box (two float64 exp) from the
//     localisation rows that arrived with the same TMA transaction as the scores, and stores the
//     box to the per-image box table (the top-k kernel gathers its winners from it);
//  C. class scan (class_scan above): lane = class, the 8 warps share the surviving valid rows.
__global__ void __launch_bounds__(kTileRows, 8)
scatter_candidates_kernel(const __grid_constant__ PostParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int C = p.C;
    const int stage_floats = (kTileRows * C + 4 + 3) & ~3;
    const int stage_bytes = stage_floats * 4 + kTileRows * 16;     // scores | loc rows
    float* const s_ninf = reinterpret_cast<float*>(smem + 2 * (size_t)stage_bytes);   // C + 1 words of -inf
    __shared__ __align__(8) u64 s_bar[2];
    __shared__ int s_rows[kTileRows];
    __shared__ int s_wcnt[kScatWarps];
    __shared__ int s_vrows[kTileRows];      // row of the j-th surviving anchor, -1 when its box failed the size test
    __shared__ int s_voff[kTileRows];       // word offset of that row's scores inside the tile (or of the -inf row)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned full = 0xffffffffu;
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        fence_proxy_async();
    }
    for (int i = tid; i <= C; i += kTileRows) s_ninf[i] = __int_as_float(0xff800000);
    __syncthreads();

    auto issue = [&](const TileInfo& t, int st) {
        // thread 0 only: one mbarrier transaction = score tile + localisation rows
        unsigned char* base = smem + (size_t)st * stage_bytes;
        unsigned cls_bytes = (unsigned)((t.a1 - t.a0) * 4);
        unsigned row_bytes = (unsigned)t.rows * 16u;
        mbar_arrive_expect_tx(&s_bar[st], cls_bytes + row_bytes);
        if (cls_bytes) tma_load_1d(base, p.cls[t.layer] + t.a0, cls_bytes, &s_bar[st]);
        const size_t row = (size_t)t.b * t.n_l + t.t0;
        tma_load_1d(base + stage_floats * 4, reinterpret_cast<const float4*>(p.loc[t.layer]) + row, row_bytes, &s_bar[st]);
    };
    // objectness of this thread's row: only the LOAD is issued here (a tile ahead); the comparison waits until
    // the value is needed, so the load's latency is hidden behind the current tile's work
    auto load_obj = [&](const TileInfo& t) -> float {
        return (p.has_obj && tid < t.rows) ? p.obj[t.layer][(size_t)t.b * t.n_l + t.t0 + tid] : 0.f;
    };

    // tiles are visited grid-stride; (b, r) = (image, tile inside the image) advances without divisions
    const int cnt = p.r_hi - p.r_lo;                    // tiles of this launch per image
    const int step_b = (int)(gridDim.x / (unsigned)cnt), step_r = (int)(gridDim.x % (unsigned)cnt);
    int b = (int)(blockIdx.x / (unsigned)cnt), r = (int)(blockIdx.x % (unsigned)cnt);
    if (b >= p.B) return;
    TileInfo t = tile_info(p, b, p.r_lo + r);
    if (tid == 0) issue(t, 0);
    float objv = load_obj(t);
    unsigned phase = 0u;   // bit st = parity the next wait on stage st must see
    for (int it = 0; b < p.B; ++it) {
        const int st = it & 1;
        // prefetch the next tile into the other stage (freed by the barrier that ended iteration it-1)
        int nb = b + step_b, nr = r + step_r;
        if (nr >= cnt) { nr -= cnt; ++nb; }
        TileInfo tn = t;
        float objv_next = 0.f;
        if (nb < p.B) {
            tn = tile_info(p, nb, p.r_lo + nr);
            if (tid == 0) issue(tn, st ^ 1);
            objv_next = load_obj(tn);           // objectness of the next tile: the load overlaps this tile's work
        }
        const bool gate = tid < t.rows && (!p.has_obj || objv > p.obj_thr);

        // ---- A. objectness gate + compaction of the surviving rows
        const unsigned gmask = __ballot_sync(full, gate);
        if (lane == 0) s_wcnt[warp] = __popc(gmask);
        __syncthreads();
        int gbase = 0, n_gated = 0;
#pragma unroll
        for (int w = 0; w < kScatWarps; ++w) {
            int c = s_wcnt[w];
            gbase += (w < warp) ? c : 0;
            n_gated += c;
        }
        if (gate) s_rows[gbase + __popc(gmask & ((1u << lane) - 1u))] = tid;
        mbar_wait(&s_bar[st], (phase >> st) & 1u);
        phase ^= 1u << st;
        __syncthreads();

        const unsigned char* base = smem + (size_t)st * stage_bytes;
        const float* s_cls = reinterpret_cast<const float*>(base) + (int)(t.start_w - t.a0);
        const float4* s_loc = reinterpret_cast<const float4*>(base + stage_floats * 4);
        const int nchunks = (n_gated + 31) >> 5;

        // ---- B. boxes of the surviving rows (compacted: thread i <-> row s_rows[i])
        if (warp < nchunks) {
            bool valid = false;
            int rr = 0;
            if (tid < n_gated) {
                rr = s_rows[tid];
                const int n = p.tab.offs[t.layer] + t.t0 + rr;
                float4 box = p.decoded ? s_loc[rr] : decode_box(s_loc[rr], p.dec[n], p.ps0, p.ps1, p.ps2, p.ps3);
                if (p.has_clip) box = clip_box(box, p.clip);
                valid = true;
                if (p.min_size >= 0.f) {
                    float h = box.z - box.x;
                    float w = box.w - box.y;
                    valid = (w > p.min_size) && (h > p.min_size);
                }
                // invalid rows get a NaN marker: the exact-rebuild path of the top-k kernel reads it
                const float qn = __int_as_float(0x7fc00000);
                p.boxes[(size_t)t.b * p.tab.N + n] = valid ? box : make_float4(qn, qn, qn, qn);
            }
            if (tid < n_gated) {
                s_vrows[tid] = valid ? rr : -1;
                s_voff[tid] = valid ? rr * C : (int)(s_ninf - s_cls);
            }
        }
        __syncthreads();

        // ---- C. class scan over the surviving rows
        if (t.end_w > t.a1)
            class_scan<true>(p, t, s_cls, s_vrows, n_gated);
        else
            class_scan_fast(p, t, s_cls, s_voff, s_vrows, n_gated);
        __syncthreads();   // stage st, s_rows and s_vrows are free again
        t = tn;
        b = nb;
        r = nr;
        objv = objv_next;
    }
}

// key source: staged prefix in shared memory, remainder straight from the global list
struct ListSrc {
    const u64* g;
    const u64* s;
    int staged;
    __device__ __forceinline__ u64 get(int i) const { return i < staged ? s[i] : g[i]; }
};

constexpr int kListCap = 2048;                               // survivors of the sampling pre-filter (shared memory)
constexpr int kKeysPerThread = kListCap / kTopkThreads;      // 8
constexpr int kSample = 2 * kTopkThreads;                    // 512 sampled keys
constexpr int kDigitBits = 11;
constexpr int kBins = 1 << kDigitBits;

// need-th largest of the keys held in registers across the block (zero = empty slot; the keys are
// distinct).  MSD radix select with 11-bit digits -- the first digit is sign + exponent + 2 mantissa
// bits of the score, so two passes almost always isolate the key -- warp-aggregated shared-memory
// histogram, block scan from the top bin.  Returns thr with #{key >= thr} == need (the low bits of
// thr are zero when a whole bucket is taken).  Requires need <= number of non-zero keys.
template <int KPT>
__device__ u64 radix_kth(const u64 (&key)[KPT], int nper, int need, unsigned* s_hist, int* s_ctl) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned full = 0xffffffffu;
    __shared__ unsigned s_wtot[kTopkThreads / 32];
    u64 prefix = 0ull, mask = 0ull;
    for (int shift = 64 - kDigitBits;; shift = max(shift - kDigitBits, 0)) {
        const int width = (shift == 0) ? (64 % kDigitBits ? 64 % kDigitBits : kDigitBits) : kDigitBits;
        const unsigned dmask = (1u << width) - 1u;
        for (int i = tid; i < kBins; i += kTopkThreads) s_hist[i] = 0u;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < KPT; ++j) {
            if (j < nper) {
                const bool act = key[j] != 0ull && (key[j] & mask) == prefix;
                const unsigned d = (unsigned)(key[j] >> shift) & dmask;
                // warp-aggregated histogram: one atomic per distinct digit per warp
                const unsigned amask = __ballot_sync(full, act);
                if (act) {
                    const unsigned peers = __match_any_sync(amask, d);
                    if (lane == __ffs(peers) - 1) atomicAdd(&s_hist[d], (unsigned)__popc(peers));
                }
            }
        }
        __syncthreads();
        // thread t owns bins [kBins-1-8t-7, kBins-1-8t] (descending); find the bin where the count of
        // keys in larger bins is < need <= that count + hist[bin]
        constexpr int BPT = kBins / kTopkThreads;
        unsigned cnt[BPT];
        unsigned sum = 0;
#pragma unroll
        for (int q = 0; q < BPT; ++q) {
            cnt[q] = s_hist[kBins - 1 - (tid * BPT + q)];
            sum += cnt[q];
        }
        unsigned incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned v = __shfl_up_sync(full, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_wtot[warp] = incl;
        __syncthreads();
        unsigned above = incl - sum;
#pragma unroll
        for (int w = 0; w < kTopkThreads / 32; ++w) above += (w < warp) ? s_wtot[w] : 0u;
        if (above < (unsigned)need && (unsigned)need <= above + sum) {
#pragma unroll
            for (int q = 0; q < BPT; ++q) {
                if (above < (unsigned)need && (unsigned)need <= above + cnt[q]) {
                    s_ctl[0] = kBins - 1 - (tid * BPT + q);
                    s_ctl[1] = need - (int)above;
                    s_ctl[2] = (cnt[q] == (unsigned)need - above) ? 1 : 0;
                    above = 0xffffffffu;   // stop
                } else if (above != 0xffffffffu) {
                    above += cnt[q];
                }
            }
        }
        __syncthreads();
        const int d = s_ctl[0];
        need = s_ctl[1];
        const int whole = s_ctl[2];
        prefix |= (u64)(unsigned)d << shift;
        mask |= (u64)dmask << shift;
        if (whole || shift == 0) break;
    }
    __syncthreads();
    return prefix;
}

// append sel keys to list (warp-aggregated: one shared atomic per warp per call)
__device__ __forceinline__ void warp_append(bool sel, u64 key, u64* list, int cap, int* counter) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned m = __ballot_sync(full, sel);
    if (m) {
        int base = 0;
        if (lane == 0) base = atomicAdd(counter, __popc(m));
        base = __shfl_sync(full, base, 0);
        if (sel) {
            const int pos = base + __popc(m & ((1u << lane) - 1u));
            if (pos < cap) list[pos] = key;
        }
    }
}

// bitonic sort of s_sort[0..P), descending, through shared memory (any power of two P)
__device__ __forceinline__ void block_bitonic_desc(u64* s_sort, int P) {
    const int tid = threadIdx.x;
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < (P >> 1); i += kTopkThreads) {
                int lo = 2 * i - (i & (stride - 1));
                int hi = lo + stride;
                bool desc = ((lo & size) == 0);
                u64 x = s_sort[lo], y = s_sort[hi];
                if ((x < y) == desc) { s_sort[lo] = y; s_sort[hi] = x; }
            }
            __syncthreads();
        }
    }
}

// Sort of s_sort[0..P), P = 256 * EPT, descending, ALL KEYS DISTINCT (empty slots hold distinct
// small values with a zero high word, see select_topk_kernel):
//  A. every warp sorts its run of 32 * EPT consecutive elements in registers (bitonic network, warp
//     shuffles, no barrier);
//  B. log2(8) merge levels: each element finds its rank in the sibling run with a branch-free binary
//     search (one dependent shared-memory read per step) and is written to its final position of the
//     merged run in the other buffer.  3 barriers per level instead of a barrier per bitonic stage,
//     and ~1/3 of the instructions.
// s_tmp: P u64 of scratch.  The result is back in s_sort.
template <int EPT>
__device__ __forceinline__ void block_sort_desc(u64* s_sort, u64* s_tmp) {
    constexpr int P = kTopkThreads * EPT;
    constexpr int L0 = 32 * EPT;
    const int tid = threadIdx.x;
    const unsigned full = 0xffffffffu;
    u64 v[EPT];
#pragma unroll
    for (int r = 0; r < EPT; ++r) v[r] = s_sort[EPT * tid + r];
#pragma unroll
    for (int size = 2; size <= L0; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= EPT) {
                const int lm = stride / EPT;
#pragma unroll
                for (int r = 0; r < EPT; ++r) {
                    const int e = (EPT * tid + r) & (L0 - 1);              // position inside the run
                    const u64 o = __shfl_xor_sync(full, v[r], lm);
                    const bool keep_max = (((e & stride) == 0) == ((e & size) == 0));
                    v[r] = (keep_max == (o > v[r])) ? o : v[r];
                }
            } else {
#pragma unroll
                for (int r = 0; r < EPT; ++r) {
                    if ((r & stride) == 0) {
                        const int e = (EPT * tid + r) & (L0 - 1);
                        const bool desc = ((e & size) == 0);
                        const u64 x = v[r], y = v[r | stride];
                        if ((x < y) == desc) { v[r] = y; v[r | stride] = x; }
                    }
                }
            }
        }
    }
    u64* src = s_sort;
    u64* dst = s_tmp;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < EPT; ++r) src[EPT * tid + r] = v[r];
    __syncthreads();
#pragma unroll
    for (int L = L0; L < P; L <<= 1) {
#pragma unroll
        for (int r = 0; r < EPT; ++r) {
            const int e = EPT * tid + r;
            const u64 x = src[e];
            const int run = e / L, i = e & (L - 1);
            const u64* sib = src + (run ^ 1) * L;
            // number of sibling elements that precede x in the merged (descending) order
            int pos = 0;
#pragma unroll
            for (int st = L >> 1; st > 0; st >>= 1)
                if (sib[pos + st - 1] > x) pos += st;
            if (sib[pos] > x) pos += 1;
            dst[(run >> 1) * 2 * L + i + pos] = x;
        }
        __syncthreads();
        u64* t = src; src = dst; dst = t;
    }
    if (src != s_sort) {
#pragma unroll
        for (int r = 0; r < EPT; ++r) s_sort[EPT * tid + r] = src[EPT * tid + r];
        __syncthreads();
    }
}

// Pivot of one (image, class) from the keys of the SAMPLED tiles (first scatter launch), one warp
// per segment: an 11-bit histogram of the score bits (exponent + 3 mantissa bits) in shared
// memory, scanned from the top; the pivot is the lower edge of the bin that holds the
// pivot_rank-th largest sampled score.  pivot_rank = expected number of top-K scores in the sample
// + 4 sigma + 8, so the second scatter launch (all other tiles, threshold = pivot) keeps every
// top-K candidate with overwhelming probability; the top-k kernel verifies that and rebuilds the
// list exactly when it did not.  Any pivot is safe; it only has to be a good guess.
constexpr int kPivotWarps = 4;
constexpr int kPivotPad = kBins + (kBins / 64) * 2;      // 2 u16 of padding per 64 bins: lane-private rows hit distinct banks
__device__ __forceinline__ int pivot_slot(unsigned d) { return (int)(d + ((d >> 6) << 1)); }

__global__ void __launch_bounds__(kPivotWarps * 32)
pivot_kernel(const __grid_constant__ PostParams p) {
    __shared__ __align__(16) unsigned short s_h[kPivotWarps][kPivotPad];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t seg = (size_t)blockIdx.x * kPivotWarps + warp;
    if (seg >= (size_t)p.B * (p.C - 1)) return;
    int n = p.counts[seg];
    n = n > p.cap ? p.cap : n;
    if (n < p.pivot_rank || n > 65535) return;            // too few samples: keep sel_thr
    unsigned short* h = s_h[warp];
    for (int i = lane; i < kPivotPad / 2; i += 32) reinterpret_cast<unsigned*>(h)[i] = 0u;
    __syncwarp();
    const u64* g = p.keys + seg * p.cap;
    for (int i0 = 0; i0 < n; i0 += 128) {
        u64 k[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = i0 + q * 32 + lane;
            k[q] = (i < n) ? g[i] : 0ull;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const bool act = k[q] != 0ull;
            const unsigned d = (unsigned)(k[q] >> 52) & (kBins - 1);     // score bits [30:20]
            const unsigned amask = __ballot_sync(full, act);
            if (act) {
                const unsigned peers = __match_any_sync(amask, d);
                if (lane == __ffs(peers) - 1) h[pivot_slot(d)] += (unsigned short)__popc(peers);   // one lane per distinct bin
            }
            __syncwarp();
        }
    }
    // lane l owns the 64 bins [64 m, 64 m + 63], m = 31 - l (descending over lanes); its 32 words are
    // contiguous and start in bank m
    const int m = 31 - lane;
    const unsigned* hw = reinterpret_cast<const unsigned*>(h) + 33 * m;
    unsigned sum = 0;
#pragma unroll 8
    for (int q = 0; q < 32; ++q) {
        const unsigned w = hw[q];
        sum += (w & 0xffffu) + (w >> 16);
    }
    unsigned incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned v = __shfl_up_sync(full, incl, o);
        if (lane >= o) incl += v;
    }
    unsigned above = incl - sum;
    const unsigned need = (unsigned)p.pivot_rank;
    if (above < need && need <= above + sum) {
        int bin = 64 * m + 63;
        for (; bin > 64 * m; --bin) {
            above += h[pivot_slot((unsigned)bin)];
            if (above >= need) break;
        }
        // scatter passes v > thr: everything in the pivot bin and above must pass
        const unsigned lower = (unsigned)bin << 20;
        const float thr = __uint_as_float(lower - 1u);
        if (thr > p.sel_thr) p.thr[seg] = thr;
    }
}

// Exact rebuild of one (image, class) list straight from the inputs (the sampled pivot turned out
// too high, probability ~1e-4 per segment): objectness gate, box validity from the box table's NaN
// marker, score > sel_thr.  Returns the new length (block-uniform).
template <int NT>
__device__ int rebuild_list(const PostParams& p, int b, int c, u64* g, int* s_cnt) {
    const int tid = threadIdx.x;
    if (tid == 0) *s_cnt = 0;
    __syncthreads();
    for (int l = 0; l < p.tab.L; ++l) {
        const int n_l = p.tab.offs[l + 1] - p.tab.offs[l];
        for (int i0 = 0; i0 < n_l; i0 += NT) {
            const int i = i0 + tid;
            bool sel = false;
            float v = 0.f;
            if (i < n_l) {
                const size_t row = (size_t)b * n_l + i;
                bool gate = true;
                if (p.has_obj) gate = p.obj[l][row] > p.obj_thr;
                if (gate) {
                    const float bx = p.boxes[(size_t)b * p.tab.N + p.tab.offs[l] + i].x;
                    v = p.cls[l][row * p.C + c];
                    sel = (bx == bx) && v > p.sel_thr;
                }
            }
            warp_append(sel, ((u64)__float_as_uint(v) << 32) | (u64)(0xffffffffu - (unsigned)(p.tab.offs[l] + i)), g,
                        p.cap, s_cnt);
        }
    }
    __syncthreads();
    const int n = *s_cnt;
    __syncthreads();          // every thread has its copy before the caller reuses the counter
    return n;
}

// One CTA per (image, class): exact top-K of the candidate list, sorted descending
// (tf.nn.top_k order: lower anchor first among equal scores), boxes gathered from the box table.
//   n <= K            every key is a winner;
//   n <= 2048         keys in registers, radix select of the K-th key;
//   longer lists      a strided sample of 512 keys gives a pivot that is below the K-th key with
//                     overwhelming probability (rank = expected + 4 sigma + 8); one streaming pass keeps
//                     the keys >= pivot in shared memory, and the select runs on those.  If the pivot
//                     turns out too high (fewer than K survivors) or too low (more than 2048), the
//                     generic 8-bit radix select over the whole list takes over -- always exact.
__global__ void __launch_bounds__(kTopkThreads, RONK_TOPK_MINB)
select_topk_kernel(const __grid_constant__ PostParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int P = next_pow2(p.K);
    u64* s_sort = reinterpret_cast<u64*>(smem);                          // P
    u64* s_list = s_sort + P;                                            // kListCap
    unsigned* s_hist = reinterpret_cast<unsigned*>(s_list + kListCap);   // kBins
    __shared__ int s_ctl[4];
    __shared__ int s_cnt;
    const int tid = threadIdx.x;
    const int seg = blockIdx.x;
    if (p.only_flagged && !p.only_flagged[seg]) return;          // second tier: only the segments that asked for more
    const int CM = p.C - 1;
    const int b = seg / CM;
    int n = p.counts[seg];
    n = n > p.cap ? p.cap : n;
    u64* g = p.keys + (size_t)seg * p.cap;
    const float seg_thr = p.thr[seg];
    for (int attempt = 0;; ++attempt) {
    // empty slots: distinct keys with a zero high word (a real key has score bits > 0 there), so the
    // sort never sees equal keys and the slots end up last
    for (int i = tid; i < P; i += kTopkThreads) s_sort[i] = (u64)(P - i);
    if (tid == 0) { s_cnt = 0; s_ctl[3] = 0; }
    __syncthreads();

    bool sorted = false;
    if (n <= p.K) {
        for (int i = tid; i < n; i += kTopkThreads) s_sort[i] = g[i];
        __syncthreads();
    } else {
        const u64* src = g;
        int m = n;
        bool generic = false;
        if (n > kListCap || (n > 4 * p.K && n > 1024)) {
            // ---- sampling pre-filter
            const float mu = (float)p.K * (float)kSample / (float)n;
            int R = (int)(mu + 4.f * sqrtf(mu) + 8.f);
            u64 pivot = 1ull;
            if (R < kSample) {
                u64 ks[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) ks[j] = g[(int)(((long long)(j * kTopkThreads + tid) * n) / kSample)];
                pivot = radix_kth<2>(ks, 2, R, s_hist, s_ctl);
            }
            for (int i0 = 0; i0 < n; i0 += kTopkThreads) {
                const int i = i0 + tid;
                u64 k = (i < n) ? g[i] : 0ull;
                warp_append(k >= pivot && k != 0ull, k, s_list, kListCap, &s_cnt);
            }
            __syncthreads();
            m = s_cnt;
            src = s_list;
            generic = (m < p.K) || (m > kListCap);
        }
        if (generic) {
            // pivot too high / too low, or K itself beyond the register path: 8-bit radix select over the whole list
            ListSrc all{g, nullptr, 0};
            block_topk_sorted(all, n, p.K, s_hist, s_ctl, s_sort);
            sorted = true;
        } else if (m <= p.K) {
            for (int i = tid; i < m; i += kTopkThreads) s_sort[i] = src[i];
            __syncthreads();
        } else {
            u64 key[kKeysPerThread];
            const int nper = (m + kTopkThreads - 1) / kTopkThreads;
#pragma unroll
            for (int j = 0; j < kKeysPerThread; ++j) {
                key[j] = 0ull;
                if (j < nper) {
                    const int i = j * kTopkThreads + tid;
                    if (i < m) key[j] = src[i];
                }
            }
            const u64 thr = radix_kth<kKeysPerThread>(key, nper, p.K, s_hist, s_ctl);
#pragma unroll
            for (int j = 0; j < kKeysPerThread; ++j)
                if (j < nper) warp_append(key[j] != 0ull && key[j] >= thr, key[j], s_sort, P, &s_ctl[3]);
            __syncthreads();
        }
    }
    if (!sorted) {
        // s_list (2048 keys) is free by now: scratch of the merge sort
        if (P == kTopkThreads) block_sort_desc<1>(s_sort, s_list);
        else if (P == 2 * kTopkThreads) block_sort_desc<2>(s_sort, s_list);
        else if (P == 4 * kTopkThreads) block_sort_desc<4>(s_sort, s_list);
        else block_bitonic_desc(s_sort, P);
    }
    // exact iff no pivot was used, or the K-th winner is above the pivot (every candidate above the
    // pivot is in the list).  Otherwise rebuild the list from the inputs and select again.
    if (attempt > 0 || !(seg_thr > p.sel_thr)) break;
    const u64 kth = s_sort[p.K - 1];
    if ((kth >> 32) != 0ull && __uint_as_float((unsigned)(kth >> 32)) > seg_thr && !p.force_rebuild) break;
    __syncthreads();
    n = rebuild_list<kTopkThreads>(p, b, seg - b * CM + 1, g, &s_cnt);
    n = n > p.cap ? p.cap : n;
    }
    for (int r = tid; r < p.K; r += kTopkThreads) {
        u64 k = s_sort[r];
        float sc = 0.f;
        float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
        int idx = -1;
        if ((k >> 32) != 0ull) {                 // real key (empty slots have a zero high word)
            sc = __uint_as_float((unsigned)(k >> 32));
            idx = (int)(0xffffffffu - (unsigned)(k & 0xffffffffull));
            box = p.boxes[(size_t)b * p.tab.N + idx];
        }
        size_t o = (size_t)seg * p.K + r;
        p.out_scores[o] = sc;
        p.out_boxes[o] = box;
        if (p.out_idx) p.out_idx[o] = idx;
    }
}

// ------------------------------------------------------------------ large top_k (crowded scenes: K in the thousands)
// One CTA of 1024 threads per (image, class).  The K winners are found by an 8-bit MSD radix select over the key list in
// global memory (L2 resident), gathered into shared memory and sorted there by a stable LSD radix sort (8-bit digits,
// ping-pong between two K-key buffers; a pass whose byte is the same for every key -- the high bytes of ~anchor -- is
// skipped) instead of a bitonic network over next_pow2(K) keys: 6 passes of K keys against 105 compare-exchange stages
// of 16 384 for K = 10 000.  Same outputs, same pivot verification / exact rebuild as select_topk_kernel.
constexpr int kLargeThreads = 1024;
constexpr int kLargeWarps = kLargeThreads / 32;

// lanes of the warp whose 8-bit digit equals this lane's (among the lanes of `among`): eight ballots instead of
// match.any, whose latency dominated the first version of this sort
__device__ __forceinline__ unsigned same_digit_lanes(unsigned d, unsigned among) {
    unsigned peers = among;
#pragma unroll
    for (int bit = 0; bit < 8; ++bit) {
        const bool one = (d >> bit) & 1u;
        const unsigned bm = __ballot_sync(0xffffffffu, one);
        peers &= one ? bm : ~bm;
    }
    return peers;
}

// K-th largest of the n distinct keys g[0..n) (n > K): MSD radix select, 8-bit digits, histogram in shared memory
__device__ u64 large_kth(const u64* g, int n, int K, unsigned* s_hist, int* s_ctl) {
    const int tid = threadIdx.x;
    u64 prefix = 0ull, mask = 0ull;
    int need = K;
    for (int shift = 56; shift >= 0; shift -= 8) {
        if (tid < 256) s_hist[tid] = 0u;
        __syncthreads();
        for (int i0 = 0; i0 < n; i0 += kLargeThreads) {
            const int i = i0 + tid;
            bool act = false;
            unsigned d = 0u;
            if (i < n) {
                const u64 k = g[i];
                act = (k & mask) == prefix;
                d = (unsigned)(k >> shift) & 255u;
            }
            const unsigned amask = __ballot_sync(0xffffffffu, act);
            if (act) {      // (match.any here: the eight-ballot form measured 7 % slower in this loop, 3 x faster in the sort)
                const unsigned peers = __match_any_sync(amask, d);
                if ((tid & 31) == __ffs(peers) - 1) atomicAdd(&s_hist[d], (unsigned)__popc(peers));
            }
        }
        __syncthreads();
        if (tid < 32) {
            // lane l owns digits [255 - 8l - 7, 255 - 8l]: find the digit d with (#keys with a larger digit) < need <= that + hist[d]
            unsigned c[8];
            unsigned sum = 0u;
#pragma unroll
            for (int q = 0; q < 8; ++q) { c[q] = s_hist[255 - (tid * 8 + q)]; sum += c[q]; }
            unsigned incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
                if (tid >= o) incl += v;
            }
            unsigned above = incl - sum;
            if (above < (unsigned)need && (unsigned)need <= incl) {
                bool done = false;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (!done && (unsigned)need <= above + c[q]) {
                        s_ctl[0] = 255 - (tid * 8 + q);
                        s_ctl[1] = need - (int)above;
                        s_ctl[2] = (c[q] == (unsigned)need - above) ? 1 : 0;
                        done = true;
                    } else if (!done) {
                        above += c[q];
                    }
                }
            }
        }
        __syncthreads();
        prefix |= (u64)(unsigned)s_ctl[0] << shift;
        mask |= 255ull << shift;
        need = s_ctl[1];
        const int whole = s_ctl[2];
        __syncthreads();
        if (whole) break;                       // every key of that bin is a winner: keys >= prefix are exactly the K largest
    }
    return prefix;
}

// stable LSD radix sort (descending) of a[0..m) on the bytes that are non-zero in `diff`; result pointer returned (a or b).
// The histogram of a pass is accumulated while the previous pass scatters (plain shared-memory atomics).
__device__ u64* large_sort_bytes(u64* a, u64* b, int m, u64 diff, unsigned* s_total, unsigned* s_base, unsigned short* s_wcnt) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned full = 0xffffffffu;
    if (m < 2 || diff == 0ull) return a;
    int byte = 0;
    while (((diff >> (8 * byte)) & 255ull) == 0ull) ++byte;
    // histogram of the first pass
    if (tid < 256) s_total[tid] = 0u;
    __syncthreads();
    for (int i = tid; i < m; i += kLargeThreads) atomicAdd(&s_total[255u - ((unsigned)(a[i] >> (8 * byte)) & 255u)], 1u);
    __syncthreads();
    while (byte < 8) {
        const int shift = 8 * byte;
        int next = byte + 1;
        while (next < 8 && ((diff >> (8 * next)) & 255ull) == 0ull) ++next;
        const int nshift = 8 * (next & 7);
        if (tid < 32) {
            // exclusive prefix over the 256 digits (8 per lane)
            unsigned c[8];
            unsigned sum = 0u;
#pragma unroll
            for (int q = 0; q < 8; ++q) { c[q] = s_total[tid * 8 + q]; sum += c[q]; }
            unsigned incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned v = __shfl_up_sync(full, incl, o);
                if (tid >= o) incl += v;
            }
            unsigned run = incl - sum;
#pragma unroll
            for (int q = 0; q < 8; ++q) { s_base[tid * 8 + q] = run; run += c[q]; s_total[tid * 8 + q] = 0u; }
        }
        reinterpret_cast<uint4*>(s_wcnt)[tid] = make_uint4(0u, 0u, 0u, 0u);            // 32 x 256 u16 = 1024 x 16 B
        __syncthreads();
        for (int i0 = 0; i0 < m; i0 += kLargeThreads) {
            const int i = i0 + tid;
            const bool valid = i < m;
            const u64 k = valid ? a[i] : 0ull;
            const unsigned d = 255u - ((unsigned)(k >> shift) & 255u);
            const unsigned peers = same_digit_lanes(d, __ballot_sync(full, valid));      // (match.any: 1.95 vs 1.70 ms per step)
            const unsigned rank = (unsigned)__popc(peers & ((1u << lane) - 1u));
            if (valid && rank == 0u) s_wcnt[warp * 256 + d] = (unsigned short)__popc(peers);
            if (valid && next < 8) atomicAdd(&s_total[255u - ((unsigned)(k >> nshift) & 255u)], 1u);
            __syncthreads();
            // exclusive prefix over the 32 warps, two digits (one 32-bit word of two u16) per thread
            unsigned run = 0u;
            if (tid < 128) {
                unsigned* col = reinterpret_cast<unsigned*>(s_wcnt) + tid;
                unsigned c[kLargeWarps];
#pragma unroll
                for (int w = 0; w < kLargeWarps; ++w) c[w] = col[w * 128];
#pragma unroll
                for (int w = 0; w < kLargeWarps; ++w) {
                    col[w * 128] = run;                    // per-half sums stay below 2^16 (<= 1024 keys per round)
                    run += c[w];
                }
            }
            __syncthreads();
            if (valid) b[s_base[d] + s_wcnt[warp * 256 + d] + rank] = k;
            __syncthreads();
            if (tid < 128) {
                s_base[2 * tid] += run & 0xffffu;
                s_base[2 * tid + 1] += run >> 16;
            }
            reinterpret_cast<uint4*>(s_wcnt)[tid] = make_uint4(0u, 0u, 0u, 0u);
            __syncthreads();
        }
        u64* t = a; a = b; b = t;
        byte = next;
    }
    return a;
}

// a[0..m) -> sorted descending (result pointer returned: a or b).  Bytes that are equal in every key (OR / AND of all
// keys) get no pass.  Scores rarely tie, so the low word (~anchor, 2 varying bytes) is left out at first: four passes over
// the score bytes, then a scan for runs of equal scores -- short runs (<= 16 keys) are put in order by the thread at their
// head, a longer run sends the segment through the full sort.
__device__ u64* large_sort_desc(u64* a, u64* b, int m, unsigned* s_total, unsigned* s_base, unsigned short* s_wcnt,
                                int* s_ctl) {
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned full = 0xffffffffu;
    // bytes that are the same in every key need no pass: OR / AND of all keys
    unsigned* s_red = reinterpret_cast<unsigned*>(s_ctl);       // or_lo, or_hi, and_lo, and_hi
    if (tid == 0) { s_red[0] = 0u; s_red[1] = 0u; s_red[2] = 0xffffffffu; s_red[3] = 0xffffffffu; }
    __syncthreads();
    {
        unsigned ol = 0u, oh = 0u, al = 0xffffffffu, ah = 0xffffffffu;
        for (int i = tid; i < m; i += kLargeThreads) {
            const u64 k = a[i];
            ol |= (unsigned)k; oh |= (unsigned)(k >> 32);
            al &= (unsigned)k; ah &= (unsigned)(k >> 32);
        }
        ol = __reduce_or_sync(full, ol); oh = __reduce_or_sync(full, oh);
        al = __reduce_and_sync(full, al); ah = __reduce_and_sync(full, ah);
        if (lane == 0) { atomicOr(&s_red[0], ol); atomicOr(&s_red[1], oh); atomicAnd(&s_red[2], al); atomicAnd(&s_red[3], ah); }
    }
    __syncthreads();
    const u64 diff = (((u64)s_red[1] << 32) | s_red[0]) ^ (((u64)s_red[3] << 32) | s_red[2]);
    __syncthreads();
    if (m < 2 || diff == 0ull) return a;
    const u64 hi = diff & 0xffffffff00000000ull;
    if (hi == 0ull || (diff & 0xffffffffull) == 0ull) return large_sort_bytes(a, b, m, diff, s_total, s_base, s_wcnt);
    u64* r = large_sort_bytes(a, b, m, hi, s_total, s_base, s_wcnt);
    u64* other = (r == a) ? b : a;
    // runs of equal scores: the heads and their lengths are found read-only and noted in the free buffer, then
    // (barrier) every head puts its own run in order -- no thread touches another run
    unsigned* runlen = reinterpret_cast<unsigned*>(other);          // m entries fit: the buffer holds K >= m keys
    if (tid == 0) s_ctl[0] = 0;
    __syncthreads();
    bool longrun = false;
    for (int i = tid; i < m; i += kLargeThreads) {
        unsigned L = 0u;
        if (i + 1 < m) {
            const unsigned sc = (unsigned)(r[i] >> 32);
            if ((unsigned)(r[i + 1] >> 32) == sc && (i == 0 || (unsigned)(r[i - 1] >> 32) != sc)) {
                L = 2u;
                while (i + (int)L < m && L <= 16u && (unsigned)(r[i + L] >> 32) == sc) ++L;
                if (L > 16u) longrun = true;
            }
        }
        runlen[i] = L;
    }
    if (longrun) s_ctl[0] = 1;
    __syncthreads();
    const bool redo = s_ctl[0] != 0;
    __syncthreads();
    if (redo) return large_sort_bytes(r, other, m, diff, s_total, s_base, s_wcnt);
    for (int i = tid; i < m; i += kLargeThreads) {
        const int L = (int)runlen[i];
        for (int x = 1; x < L; ++x) {                                            // insertion sort, descending keys
            const u64 k = r[i + x];
            int y = x - 1;
            while (y >= 0 && r[i + y] < k) { r[i + y + 1] = r[i + y]; --y; }
            r[i + y + 1] = k;
        }
    }
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kLargeThreads, 1)
select_topk_large_kernel(const __grid_constant__ PostParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    u64* s_a = reinterpret_cast<u64*>(smem);                                   // K
    u64* s_b = s_a + p.K;                                                      // K
    unsigned short* s_wcnt = reinterpret_cast<unsigned short*>(s_b + p.K);      // 32 warps x 256 digits
    unsigned* s_total = reinterpret_cast<unsigned*>(s_wcnt + kLargeWarps * 256);
    unsigned* s_base = s_total + 256;
    __shared__ int s_ctl[4];
    __shared__ int s_cnt;
    const int tid = threadIdx.x;
    const int seg = blockIdx.x;
    if (p.only_flagged && !p.only_flagged[seg]) return;
    const int CM = p.C - 1;
    const int b = seg / CM;
    int n = p.counts[seg];
    n = n > p.cap ? p.cap : n;
    u64* g = p.keys + (size_t)seg * p.cap;
    const float seg_thr = p.thr[seg];
    const u64* sorted = s_a;
    int m = 0;
    for (int attempt = 0;; ++attempt) {
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        if (n <= p.K) {
            for (int i = tid; i < n; i += kLargeThreads) {
                const u64 k = g[i];
                s_a[i] = k;
            }
            m = n;
        } else if (n <= 2 * p.K) {
            // the whole list fits in the two sort buffers: one pass over global memory, the select passes read shared
            // memory, and the winners are compacted in place -- after round r at most 1024 (r + 1) winners exist, so a
            // write never reaches a key that has not been read yet (reads of a round complete before its writes)
            for (int i = tid; i < n; i += kLargeThreads) s_a[i] = g[i];
            __syncthreads();
            const u64 thr = large_kth(s_a, n, p.K, s_total, s_ctl);
            for (int i0 = 0; i0 < n; i0 += kLargeThreads) {
                const int i = i0 + tid;
                const u64 k = (i < n) ? s_a[i] : 0ull;
                __syncthreads();
                warp_append(k >= thr && k != 0ull, k, s_a, p.K, &s_cnt);
                __syncthreads();
            }
            m = p.K;
        } else {
            const u64 thr = large_kth(g, n, p.K, s_total, s_ctl);
            for (int i0 = 0; i0 < n; i0 += kLargeThreads) {
                const int i = i0 + tid;
                const u64 k = (i < n) ? g[i] : 0ull;
                warp_append(k >= thr && k != 0ull, k, s_a, p.K, &s_cnt);
            }
            m = p.K;
        }
        __syncthreads();
        sorted = large_sort_desc(s_a, s_b, m, s_total, s_base, s_wcnt, s_ctl);
        // exact iff no pivot was used, or the K-th winner lies above the pivot (see select_topk_kernel)
        if (attempt > 0 || !(seg_thr > p.sel_thr)) break;
        if (m >= p.K && __uint_as_float((unsigned)(sorted[p.K - 1] >> 32)) > seg_thr && !p.force_rebuild) break;
        __syncthreads();
        n = rebuild_list<kLargeThreads>(p, b, seg - b * CM + 1, g, &s_cnt);
        n = n > p.cap ? p.cap : n;
    }
    for (int r = tid; r < p.K; r += kLargeThreads) {
        float sc = 0.f;
        float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
        int idx = -1;
        if (r < m) {
            const u64 k = sorted[r];
            sc = __uint_as_float((unsigned)(k >> 32));
            idx = (int)(0xffffffffu - (unsigned)(k & 0xffffffffull));
            box = p.boxes[(size_t)b * p.tab.N + idx];
        }
        const size_t o = (size_t)seg * p.K + r;
        p.out_scores[o] = sc;
        p.out_boxes[o] = box;
        if (p.out_idx) p.out_idx[o] = idx;
    }
}

__global__ void init_segments_kernel(int* counts, float* thr, float sel_thr, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        counts[i] = 0;
        thr[i] = sel_thr;
    }
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
    unsigned u = __float_as_uint(s + 0.f);          // -0 -> +0: tf.nn.top_k compares values, the two zeros tie
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
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
            idx = (int)(0xffffffffu - (unsigned)(k & 0xffffffffull));
            sc = scores[row * N + idx];
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
    return 2 * align_up(segs * sizeof(int), 256) + segs * (size_t)h->tab.N * sizeof(u64) + (size_t)B * h->tab.N * 16;
}

// argument checks + everything of PostParams that does not depend on which launches follow
static int fill_post_params(PostParams& p, const char* who, const ronk_anchors_t* h, const float* const* loc_layers,
                            const float* const* cls_layers, const float* const* obj_layers, int B, int C,
                            float objectness_threshold, float select_threshold, const float* clip, float min_size,
                            const float* ps, int K, int select_flags, float* out_scores, float* out_boxes, int32_t* out_idx,
                            void* ws) {
    (void)who;
    RONK_REQUIRE(h != nullptr, RONK_EINVAL, "ronk_decode_select_topk: NULL anchor handle");
    RONK_REQUIRE(loc_layers && cls_layers && ps && out_scores && out_boxes && ws, RONK_EINVAL,
                 "ronk_decode_select_topk: NULL pointer argument");
    RONK_REQUIRE(B >= 1 && C >= 2 && C <= 1024, RONK_EINVAL, "ronk_decode_select_topk: need B >= 1 and 2 <= C <= 1024");
    RONK_REQUIRE(K >= 1 && K <= 16384, RONK_ELIMIT, "ronk_decode_select_topk: need 1 <= K <= 16384");
    RONK_REQUIRE(select_threshold >= 0.f, RONK_EINVAL,
                 "ronk_decode_select_topk: select_threshold must be >= 0 (None in the reference is 0)");
    RONK_REQUIRE(((uintptr_t)out_boxes % 16) == 0 && ((uintptr_t)ws % 256) == 0, RONK_EINVAL,
                 "ronk_decode_select_topk: out_boxes must be 16-byte and ws 256-byte aligned");
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
    p.thr = (float*)((char*)ws + align_up(segs * sizeof(int), 256));
    p.keys = (u64*)((char*)ws + 2 * align_up(segs * sizeof(int), 256));
    p.boxes = (float4*)(p.keys + segs * (size_t)h->tab.N);
    p.cap = h->tab.N;
    p.num_tiles = (long long)toff * B;
    RONK_REQUIRE(p.num_tiles < (1ll << 31), RONK_ELIMIT, "ronk_decode_select_topk: too many tiles (B * anchors)");
    p.out_scores = out_scores;
    p.out_boxes = (float4*)out_boxes;
    p.out_idx = out_idx;
    p.only_flagged = nullptr;
    p.force_rebuild = (select_flags & RONK_SELECT_TEST_REBUILD) ? 1 : 0;
    p.pivot_rank = 0;
    p.tile_tab = (const int4*)h->d_tile_tab;
    p.r_lo = p.r_hi = 0;
    return RONK_OK;
}

static int launch_select_topk(const PostParams& p, size_t segs, cudaStream_t st) {
    // K beyond the register / rank-merge paths of select_topk_kernel (its fallback is a 16 384-key bitonic network):
    // the radix-sort kernel, as long as two K-key buffers fit in shared memory.  RONK_TOPK_LARGE=0 keeps the old path.
    const size_t smem_l = (size_t)2 * p.K * sizeof(u64) + (size_t)kLargeWarps * 256 * 2 + 512 * sizeof(unsigned);
    bool large = p.K > 4 * kTopkThreads && smem_l <= 220 * 1024;
    if (const char* e = getenv("RONK_TOPK_LARGE")) large = large && e[0] != '0';
    if (large) {
        if (smem_l > 48 * 1024)
            RONK_CUDA(cudaFuncSetAttribute(select_topk_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l));
        select_topk_large_kernel<<<(unsigned)segs, kLargeThreads, smem_l, st>>>(p);
        RONK_LAUNCHED();
        return RONK_OK;
    }
    int P = 1;
    while (P < p.K) P <<= 1;
    size_t smem_b = (size_t)(P + kListCap) * sizeof(u64) + (size_t)kBins * sizeof(unsigned);
    if (smem_b > 48 * 1024)
        RONK_CUDA(cudaFuncSetAttribute(select_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
    select_topk_kernel<<<(unsigned)segs, kTopkThreads, smem_b, st>>>(p);
    RONK_LAUNCHED();
    return RONK_OK;
}

// Second tier of a two-tier top-k (see ronk.h): the key lists / box table of an earlier ronk_decode_select_topk call
// on the same inputs and workspace are selected again with a larger K, for the segments whose flag is set only.
extern "C" int ronk_select_topk_flagged(const ronk_anchors_t* h, const float* const* loc_layers,
                                        const float* const* cls_layers, const float* const* obj_layers, int B, int C,
                                        float objectness_threshold, float select_threshold, const float* clip,
                                        float min_size, const float* ps, int K, int select_flags, const int32_t* flags,
                                        float* out_scores, float* out_boxes, int32_t* out_idx, void* ws, void* stream) {
    RONK_REQUIRE(flags != nullptr, RONK_EINVAL, "ronk_select_topk_flagged: NULL flags");
    PostParams p;
    if (int rc = fill_post_params(p, "ronk_select_topk_flagged", h, loc_layers, cls_layers, obj_layers, B, C, objectness_threshold,
                                  select_threshold, clip, min_size, ps, K, select_flags, out_scores, out_boxes, out_idx, ws))
        return rc;
    p.only_flagged = flags;
    return launch_select_topk(p, (size_t)B * (C - 1), (cudaStream_t)stream);
}

extern "C" int ronk_decode_select_topk(const ronk_anchors_t* h, const float* const* loc_layers,
                                       const float* const* cls_layers, const float* const* obj_layers, int B,
                                       int C, float objectness_threshold, float select_threshold,
                                       const float* clip, float min_size, const float* ps, int K,
                                       int select_flags, float* out_scores, float* out_boxes, int32_t* out_idx, void* ws,
                                       void* stream) {
    PostParams p;
    if (int rc = fill_post_params(p, "ronk_decode_select_topk", h, loc_layers, cls_layers, obj_layers, B, C, objectness_threshold,
                                  select_threshold, clip, min_size, ps, K, select_flags, out_scores, out_boxes, out_idx, ws))
        return rc;
    const size_t segs = (size_t)B * (C - 1);
    const int toff = p.tile_off[h->tab.L];
    cudaStream_t st = (cudaStream_t)stream;

    init_segments_kernel<<<(unsigned)((segs + 255) / 256), 256, 0, st>>>(p.counts, p.thr, p.sel_thr, segs);
    RONK_LAUNCHED();

    const int stage_floats = (kTileRows * C + 4 + 3) & ~3;
    size_t smem_a = (size_t)2 * ((size_t)stage_floats * 4 + kTileRows * 16) + (size_t)(C + 1 + 3) / 4 * 16;
    RONK_REQUIRE(smem_a <= 220 * 1024, RONK_ELIMIT, "ronk_decode_select_topk: C too large for the shared-memory tile");
    if (smem_a > 48 * 1024)
        RONK_CUDA(cudaFuncSetAttribute(scatter_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
    int per_sm = (int)((220 * 1024) / (smem_a + 2048));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    // Two scatter launches when the lists would be much longer than K: first a spread-out sixteenth of
    // every image's tiles (threshold sel_thr), then a per-(image, class) pivot from those keys,
    // then all other tiles with the pivot as threshold.  perm_mul ~ 0.618 * tiles, coprime.
    const int tpi = h->tiles_per_image;        // == toff; the handle's tile table is in the permuted order
    const int m = h->tile_perm_mul;
    int tpi1 = 0;
    if (!(select_flags & RONK_SELECT_NO_SAMPLING) && tpi >= 16 && (long long)K * 8 <= h->tab.N) {
        tpi1 = tpi / 16;                 // a sixteenth of the tiles: pivot rank mu + 4 sigma + 8 still leaves ~2K keys per list
        if (tpi1 < 2) tpi1 = 2;
        if (tpi1 > 16) tpi1 = 16;
        long long rows = 0;
        for (int q = 0; q < tpi1; ++q) {
            int r = (int)(((long long)q * m) % tpi);
            int l = 0;
            while (l + 1 < h->tab.L && r >= p.tile_off[l + 1]) ++l;
            int n_l = h->tab.offs[l + 1] - h->tab.offs[l];
            int t0 = (r - p.tile_off[l]) * kTileRows;
            rows += (n_l - t0 < kTileRows) ? (n_l - t0) : kTileRows;
        }
        const double mu = (double)K * (double)rows / (double)h->tab.N;
        p.pivot_rank = (int)(mu + 4.0 * sqrt(mu) + 8.0);
        if (p.force_rebuild) p.pivot_rank = 1;       // pivot = the highest sampled score: (almost) always too high
    }
    auto launch_scatter = [&](int r_lo, int r_hi) -> int {
        p.r_lo = r_lo;
        p.r_hi = r_hi;
        long long tiles = (long long)(r_hi - r_lo) * B;
        long long grid = (long long)h->num_sms * per_sm;
        if (grid > tiles) grid = tiles;
        scatter_candidates_kernel<<<(unsigned)grid, kTileRows, smem_a, st>>>(p);
        RONK_LAUNCHED();
        return RONK_OK;
    };
    if (tpi1 > 0) {
        if (int rc = launch_scatter(0, tpi1)) return rc;
        pivot_kernel<<<(unsigned)((segs + kPivotWarps - 1) / kPivotWarps), kPivotWarps * 32, 0, st>>>(p);
        RONK_LAUNCHED();
        if (int rc = launch_scatter(tpi1, tpi)) return rc;
    } else {
        if (int rc = launch_scatter(0, tpi)) return rc;
    }

    return launch_select_topk(p, segs, st);
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
