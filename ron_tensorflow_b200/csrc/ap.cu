// Average precision straight from the device-resident TP/FP records: the tail of the evaluation loop
// (tf_extended/metrics.py:100-130 precision_recall, :212-234 average_precision_voc12, :237-258
// average_precision_voc07; called from eval_ron_network.py:262-324) without moving the records to the host.
//
//   records  uint64 [n]: score bits << 32 | class index << 8 | fp << 1 | tp   (ronk_tpfp_records_append), in the order
//            the reference would have concatenated them (rank-major, batch after batch); class index 0xffffff marks a
//            padding entry (the unused tail of an all-gathered row) and is ignored.
//
// 1. Stable LSD radix sort (8-bit digits) by (class ascending, score descending): tf.nn.top_k(sorted=True) keeps the
//    lower index first among equal scores, and a stable sort of records in concatenation order does exactly that.
// 2. Per class (grid.y = class): cumulative TP / FP counts (exact integers), precision = tp / (tp + fp) and
//    recall = tp / n_gt in float64 with the reference's safe division, reverse running maximum of the precision (the
//    interpolated envelope), the VOC12 Riemann sum and the 11 VOC07 look-ups (recall is non-decreasing: the set
//    recall >= t is a suffix, found by binary search).  Every reduction has a fixed order: results are reproducible
//    bit for bit, VOC07 equals the host / oracle value exactly, VOC12 differs from NumPy's pairwise sum by rounding
//    (~1e-16 relative).
#include "common.cuh"

namespace ronk {

constexpr int kSortThreads = 256;
constexpr int kSortRounds = 32;
constexpr int kSortTile = kSortThreads * kSortRounds;      // 8192 records per CTA

__device__ __forceinline__ unsigned rec_class(u64 r) { return (unsigned)(r & 0xffffffffull) >> 8; }

// records (pass < 8): passes 0..3 = bytes of the inverted score bits (descending score), 4..6 = bytes of the class
// index.  Plain keys (pass >= 8, ronk_sort_rows): byte (pass - 8) of the key, ascending.
__device__ __forceinline__ unsigned sort_digit(u64 r, int pass) {
    if (pass >= 8) return (unsigned)(r >> (8 * (pass - 8))) & 255u;
    if (pass < 4) return ((~(unsigned)(r >> 32)) >> (8 * pass)) & 255u;
    return (rec_class(r) >> (8 * (pass - 4))) & 255u;
}

// histogram of one digit per CTA tile, stored digit-major: hist[digit * nblocks + block]
__global__ void __launch_bounds__(kSortThreads)
ap_sort_hist_kernel(const u64* __restrict__ in, long long n, int pass, unsigned* __restrict__ hist, int nblocks) {
    __shared__ unsigned s_h[256];
    const int tid = threadIdx.x;
    s_h[tid] = 0u;
    __syncthreads();
    const long long base = (long long)blockIdx.x * kSortTile;
#pragma unroll 4
    for (int k = 0; k < kSortRounds; ++k) {
        const long long i = base + (long long)k * kSortThreads + tid;
        const bool valid = i < n;
        const unsigned d = valid ? sort_digit(in[i], pass) : 0xffffffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (valid && (tid & 31) == __ffs(peers) - 1) atomicAdd(&s_h[d], (unsigned)__popc(peers));
    }
    __syncthreads();
    hist[(size_t)tid * nblocks + blockIdx.x] = s_h[tid];
}

// exclusive prefix sum of a[0..L) in place, one CTA of 1024 threads, 16 consecutive entries per thread and step
__global__ void __launch_bounds__(1024)
ap_scan_u32_kernel(unsigned* a, int L) {
    __shared__ unsigned s_w[32];
    __shared__ unsigned s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned carry = 0u;
    for (int base = 0; base < L; base += 1024 * 16) {
        const int i0 = base + tid * 16;
        unsigned v[16];
        unsigned sum = 0u;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            v[k] = (i0 + k < L) ? a[i0 + k] : 0u;
            sum += v[k];
        }
        unsigned incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned w = s_w[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned u = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += u;
            }
            s_w[lane] = wi - w;
            if (lane == 31) s_total = wi;
        }
        __syncthreads();
        unsigned run = carry + s_w[warp] + (incl - sum);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (i0 + k < L) a[i0 + k] = run;
            run += v[k];
        }
        carry += s_total;
        __syncthreads();
    }
}

// stable scatter of one pass: rounds of 256 consecutive records; inside a round the rank of a record among the
// equal digits before it = (records of the earlier warps) + (earlier lanes of its own warp, match_any)
__global__ void __launch_bounds__(kSortThreads)
ap_sort_scatter_kernel(const u64* __restrict__ in, u64* __restrict__ out, long long n, int pass,
                       const unsigned* __restrict__ hist, int nblocks) {
    __shared__ unsigned s_base[256];
    __shared__ unsigned s_wcnt[kSortThreads / 32][256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    s_base[tid] = hist[(size_t)tid * nblocks + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kSortThreads / 32; ++w) s_wcnt[w][tid] = 0u;
    __syncthreads();
    const long long base = (long long)blockIdx.x * kSortTile;
#pragma unroll 1
    for (int k = 0; k < kSortRounds; ++k) {
        const long long i = base + (long long)k * kSortThreads + tid;
        if (base + (long long)k * kSortThreads >= n) break;             // block-uniform
        const bool valid = i < n;
        const u64 r = valid ? in[i] : 0ull;
        const unsigned d = valid ? sort_digit(r, pass) : 0xffffffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const unsigned rank = (unsigned)__popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0u) s_wcnt[warp][d] = (unsigned)__popc(peers);
        __syncthreads();
        if (valid) {
            unsigned pos = s_base[d] + rank;
            for (int w = 0; w < warp; ++w) pos += s_wcnt[w][d];
            out[pos] = r;
        }
        __syncthreads();
        unsigned add = 0u;
#pragma unroll
        for (int w = 0; w < kSortThreads / 32; ++w) {
            add += s_wcnt[w][tid];
            s_wcnt[w][tid] = 0u;
        }
        s_base[tid] += add;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
constexpr int kApThreads = 256;
constexpr int kApItems = 8;
constexpr int kApTile = kApThreads * kApItems;     // 2048 records of one class per CTA
constexpr int kApMaxThr = 16;

struct ApParams {
    const u64* rec;          // sorted records
    long long n;
    int CM, T;               // classes, tiles per class (upper bound: ceil(n / kApTile), >= 1)
    const long long* n_gt;   // [CM]
    int* off;                // [CM + 1] first record of every class; off[CM] = number of real records
    uint2* part;             // [CM * T] (tp, fp) per tile -> exclusive prefix inside the class
    double* pmax;            // [CM * T] largest precision of the tile -> largest of the LATER tiles
    double* part12;          // [CM * T] VOC12 partial sums
    double* prec;            // [n]
    unsigned* ctp;           // [n] cumulative true positives
    int* sidx;               // [CM * kApMaxThr] first record (inside the class) whose recall reaches the threshold
    double* v07;             // [CM * kApMaxThr]
    int n_thr;
    double thr[kApMaxThr];
    double* ap07;            // [CM]
    double* ap12;            // [CM]
    double* out_precision;   // [n] or NULL (sorted order)
    double* out_recall;      // [n] or NULL
};

__global__ void ap_offsets_kernel(const __grid_constant__ ApParams p) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > p.CM) return;
    long long lo = 0, hi = p.n;                     // first index whose class >= k
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (rec_class(p.rec[mid]) >= (unsigned)k) hi = mid; else lo = mid + 1;
    }
    p.off[k] = (int)lo;
}

__device__ __forceinline__ bool ap_tile(const ApParams& p, int* c, int* lo, int* len, int* t0) {
    *c = blockIdx.y;
    *lo = p.off[*c];
    *len = p.off[*c + 1] - *lo;
    *t0 = blockIdx.x * kApTile;
    return *t0 < *len;
}

// (tp, fp) of every tile
__global__ void __launch_bounds__(kApThreads)
ap_partial_kernel(const __grid_constant__ ApParams p) {
    __shared__ unsigned s_t[kApThreads / 32], s_f[kApThreads / 32];
    int c, lo, len, t0;
    if (!ap_tile(p, &c, &lo, &len, &t0)) return;
    const int tid = threadIdx.x;
    unsigned t = 0u, f = 0u;
#pragma unroll
    for (int k = 0; k < kApItems; ++k) {
        const int i = t0 + k * kApThreads + tid;
        if (i < len) {
            const unsigned m = (unsigned)p.rec[lo + i];
            t += m & 1u;
            f += (m >> 1) & 1u;
        }
    }
    t = __reduce_add_sync(0xffffffffu, t);
    f = __reduce_add_sync(0xffffffffu, f);
    if ((tid & 31) == 0) { s_t[tid >> 5] = t; s_f[tid >> 5] = f; }
    __syncthreads();
    if (tid == 0) {
        unsigned a = 0u, b = 0u;
        for (int w = 0; w < kApThreads / 32; ++w) { a += s_t[w]; b += s_f[w]; }
        p.part[(size_t)c * p.T + blockIdx.x] = make_uint2(a, b);
    }
}

// per class: exclusive prefix of the tile sums (one CTA per class, chunks of 256 tiles with a carry)
__global__ void __launch_bounds__(kApThreads)
ap_partial_scan_kernel(const __grid_constant__ ApParams p) {
    __shared__ unsigned s_t[kApThreads / 32], s_f[kApThreads / 32];
    __shared__ unsigned s_tot[2];
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int len = p.off[c + 1] - p.off[c];
    const int tiles = (len + kApTile - 1) / kApTile;
    uint2* a = p.part + (size_t)c * p.T;
    unsigned ct = 0u, cf = 0u;
    for (int base = 0; base < tiles; base += kApThreads) {
        const int i = base + tid;
        const uint2 v = (i < tiles) ? a[i] : make_uint2(0u, 0u);
        unsigned it = v.x, jf = v.y;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned x = __shfl_up_sync(0xffffffffu, it, o), y = __shfl_up_sync(0xffffffffu, jf, o);
            if (lane >= o) { it += x; jf += y; }
        }
        if (lane == 31) { s_t[warp] = it; s_f[warp] = jf; }
        __syncthreads();
        unsigned bt = 0u, bf = 0u;
        for (int w = 0; w < warp; ++w) { bt += s_t[w]; bf += s_f[w]; }
        if (tid == kApThreads - 1) { s_tot[0] = bt + it; s_tot[1] = bf + jf; }
        if (i < tiles) a[i] = make_uint2(ct + bt + it - v.x, cf + bf + jf - v.y);
        __syncthreads();
        ct += s_tot[0];
        cf += s_tot[1];
        __syncthreads();
    }
}

__device__ __forceinline__ double ap_recall(unsigned ctp, long long ng) {
    return ng > 0 ? (double)ctp / (double)ng : 0.0;            // _safe_div (metrics.py:128)
}

// cumulative counts + precision of every record (thread = kApItems CONSECUTIVE records), largest precision of the tile
__global__ void __launch_bounds__(kApThreads)
ap_precision_kernel(const __grid_constant__ ApParams p) {
    __shared__ unsigned s_t[kApThreads / 32], s_f[kApThreads / 32];
    __shared__ double s_m[kApThreads / 32];
    int c, lo, len, t0;
    if (!ap_tile(p, &c, &lo, &len, &t0)) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i0 = t0 + tid * kApItems;
    unsigned m[kApItems];
    unsigned t = 0u, f = 0u;
#pragma unroll
    for (int k = 0; k < kApItems; ++k) {
        m[k] = (i0 + k < len) ? (unsigned)p.rec[lo + i0 + k] & 3u : 0u;
        t += m[k] & 1u;
        f += m[k] >> 1;
    }
    unsigned it = t, jf = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned x = __shfl_up_sync(0xffffffffu, it, o), y = __shfl_up_sync(0xffffffffu, jf, o);
        if (lane >= o) { it += x; jf += y; }
    }
    if (lane == 31) { s_t[warp] = it; s_f[warp] = jf; }
    __syncthreads();
    const uint2 before = p.part[(size_t)c * p.T + blockIdx.x];
    unsigned ct = before.x + it - t, cf = before.y + jf - f;
    for (int w = 0; w < warp; ++w) { ct += s_t[w]; cf += s_f[w]; }
    const long long ng = p.n_gt[c];
    double best = 0.0;
#pragma unroll
    for (int k = 0; k < kApItems; ++k) {
        if (i0 + k < len) {
            ct += m[k] & 1u;
            cf += m[k] >> 1;
            const double den = (double)ct + (double)cf;
            const double pr = den > 0.0 ? (double)ct / den : 0.0;       // _safe_div (metrics.py:129)
            p.prec[lo + i0 + k] = pr;
            p.ctp[lo + i0 + k] = ct;
            if (p.out_precision) p.out_precision[lo + i0 + k] = pr;
            if (p.out_recall) p.out_recall[lo + i0 + k] = ap_recall(ct, ng);
            best = fmax(best, pr);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) s_m[warp] = best;
    __syncthreads();
    if (tid == 0) {
        double b = 0.0;
        for (int w = 0; w < kApThreads / 32; ++w) b = fmax(b, s_m[w]);
        p.pmax[(size_t)c * p.T + blockIdx.x] = b;
    }
}

// per class: pmax[tile] <- largest precision of the LATER tiles (0 behind the last one: the appended [0.] of
// metrics.py:224 / :248), and the VOC07 look-up positions
__global__ void __launch_bounds__(kApThreads)
ap_suffix_kernel(const __grid_constant__ ApParams p) {
    __shared__ double s_w[kApThreads / 32];
    __shared__ double s_tot;
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lo = p.off[c], len = p.off[c + 1] - lo;
    const int tiles = (len + kApTile - 1) / kApTile;
    double* a = p.pmax + (size_t)c * p.T;
    double carry = 0.0;
    // chunks from the end; thread j of a chunk holds tile (end - 1 - j): a forward scan over j is a suffix scan over tiles
    for (int end = tiles; end > 0; end -= kApThreads) {
        const int i = end - 1 - tid;
        const double v = (i >= 0) ? a[i] : 0.0;
        double inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double x = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc = fmax(inc, x);
        }
        double excl = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) excl = 0.0;
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        double b = carry;
        for (int w = 0; w < warp; ++w) b = fmax(b, s_w[w]);
        if (tid == kApThreads - 1) s_tot = fmax(b, inc);
        if (i >= 0) a[i] = fmax(b, excl);
        __syncthreads();
        carry = s_tot;
        __syncthreads();
    }
    if (tid < p.n_thr) {
        const long long ng = p.n_gt[c];
        const double t = p.thr[tid];
        int l = 0, h = len;                                   // first record with recall >= t (recall never decreases)
        while (l < h) {
            const int mid = (l + h) >> 1;
            if (ap_recall(p.ctp[lo + mid], ng) >= t) h = mid; else l = mid + 1;
        }
        p.sidx[c * kApMaxThr + tid] = l;
        p.v07[c * kApMaxThr + tid] = 0.0;                     // the appended (precision 0, recall inf) entry
    }
}

// envelope (reverse running maximum of the precision), VOC12 terms, VOC07 look-ups
__global__ void __launch_bounds__(kApThreads)
ap_envelope_kernel(const __grid_constant__ ApParams p) {
    __shared__ double s_w[kApThreads / 32];
    __shared__ double s_sum[kApThreads / 32];
    __shared__ int s_idx[kApMaxThr];
    int c, lo, len, t0;
    if (!ap_tile(p, &c, &lo, &len, &t0)) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < kApMaxThr) s_idx[tid] = tid < p.n_thr ? p.sidx[c * kApMaxThr + tid] : -1;
    const int i0 = t0 + tid * kApItems;
    double pr[kApItems];
    double tmax = 0.0;
#pragma unroll
    for (int k = 0; k < kApItems; ++k) {
        pr[k] = (i0 + k < len) ? p.prec[lo + i0 + k] : 0.0;
        tmax = fmax(tmax, pr[k]);
    }
    // suffix maximum over the LATER threads of the tile: a forward scan in mirrored lane / warp order
    double inc = tmax;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double x = __shfl_down_sync(0xffffffffu, inc, o);
        if (lane + o < 32) inc = fmax(inc, x);
    }
    double later = __shfl_down_sync(0xffffffffu, inc, 1);
    if (lane == 31) later = 0.0;
    if (lane == 0) s_w[warp] = inc;
    __syncthreads();
    double env = fmax(later, p.pmax[(size_t)c * p.T + blockIdx.x]);
    for (int w = warp + 1; w < kApThreads / 32; ++w) env = fmax(env, s_w[w]);
    const long long ng = p.n_gt[c];
    double e[kApItems];
#pragma unroll
    for (int k = kApItems - 1; k >= 0; --k) {
        env = fmax(env, pr[k]);
        e[k] = env;
    }
    double rprev = (i0 > 0 && i0 - 1 < len) ? ap_recall(p.ctp[lo + i0 - 1], ng) : 0.0;     // the prepended [0.] (metrics.py:225)
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < kApItems; ++k) {
        if (i0 + k < len) {
            const double r = ap_recall(p.ctp[lo + i0 + k], ng);
            const double term = e[k] * (r - rprev);           // mean_pre * diff_rec (metrics.py:231-233)
            sum = sum + term;
            rprev = r;
#pragma unroll 1
            for (int q = 0; q < p.n_thr; ++q)
                if (s_idx[q] == i0 + k) p.v07[c * kApMaxThr + q] = e[k];
        }
    }
    // fixed-order block sum
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum = sum + __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) s_sum[warp] = sum;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < kApThreads / 32; ++w) s = s + s_sum[w];
        p.part12[(size_t)c * p.T + blockIdx.x] = s;
    }
}

__global__ void ap_final_kernel(const __grid_constant__ ApParams p) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.CM) return;
    const int len = p.off[c + 1] - p.off[c];
    const int tiles = (len + kApTile - 1) / kApTile;
    double s = 0.0;
    for (int t = 0; t < tiles; ++t) s = s + p.part12[(size_t)c * p.T + t];
    p.ap12[c] = s;                                            // the last term is 0 * (1 - recall[-1])
    double a = 0.0;
    for (int q = 0; q < p.n_thr; ++q) a = a + p.v07[c * kApMaxThr + q] / (double)p.n_thr;     // l_aps.append(v / 11.) ; add_n
    p.ap07[c] = a;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct ApLayout {
    size_t buf0, buf1, hist, off, part, pmax, part12, prec, ctp, sidx, v07, total;
    int nblocks, T;
};

static ApLayout ap_layout(long long n, int CM) {
    ApLayout L;
    const long long n1 = n > 0 ? n : 1;
    L.nblocks = (int)((n1 + kSortTile - 1) / kSortTile);
    L.T = (int)((n1 + kApTile - 1) / kApTile);
    size_t o = 0;
    L.buf0 = o; o += align256((size_t)n1 * 8);
    L.buf1 = o; o += align256((size_t)n1 * 8);
    L.hist = o; o += align256((size_t)256 * L.nblocks * 4);
    L.off = o; o += align256((size_t)(CM + 1) * 4);
    L.part = o; o += align256((size_t)CM * L.T * 8);
    L.pmax = o; o += align256((size_t)CM * L.T * 8);
    L.part12 = o; o += align256((size_t)CM * L.T * 8);
    L.prec = o; o += align256((size_t)n1 * 8);
    L.ctp = o; o += align256((size_t)n1 * 4);
    L.sidx = o; o += align256((size_t)CM * kApMaxThr * 4);
    L.v07 = o; o += align256((size_t)CM * kApMaxThr * 8);
    L.total = o;
    return L;
}

// one stable pass of the LSD radix sort: src -> dst by digit `pass`
static int radix_pass(const u64* src, u64* dst, long long n, int pass, unsigned* hist, int nblocks, cudaStream_t st) {
    ap_sort_hist_kernel<<<nblocks, kSortThreads, 0, st>>>(src, n, pass, hist, nblocks);
    RONK_LAUNCHED();
    ap_scan_u32_kernel<<<1, 1024, 0, st>>>(hist, 256 * nblocks);
    RONK_LAUNCHED();
    ap_sort_scatter_kernel<<<nblocks, kSortThreads, 0, st>>>(src, dst, n, pass, hist, nblocks);
    RONK_LAUNCHED();
    return RONK_OK;
}

// ------------------------------------------------------------------ rows longer than the shared-memory sort takes
// ronk_sort_rows: tf.nn.top_k(sorted=True) of every row of [S,N] for any N < 2^24: one composite key per element,
// row << 56 | (0xffffffff - orderable(score)) << 24 | column, sorted ascending by the bytes that can differ.
__device__ __forceinline__ unsigned rows_orderable(float s) {
    const unsigned u = __float_as_uint(s + 0.f);          // -0 -> +0: the two zeros tie (lower index first)
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(256)
rows_keys_kernel(const float* __restrict__ scores, int N, long long total, u64* __restrict__ keys) {
    const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
    if (e >= total) return;
    const unsigned row = (unsigned)(e / N), col = (unsigned)(e - (long long)row * N);
    keys[e] = ((u64)row << 56) | ((u64)(0xffffffffu - rows_orderable(scores[e])) << 24) | (u64)col;
}

__global__ void __launch_bounds__(256)
rows_gather_kernel(const u64* __restrict__ keys, const float* __restrict__ scores, const float4* __restrict__ boxes, int N,
                   int K, long long total, float* __restrict__ out_scores, float4* __restrict__ out_boxes,
                   int* __restrict__ out_idx) {
    const long long e = (long long)blockIdx.x * 256 + threadIdx.x;          // (row, rank) of the output
    if (e >= total) return;
    const int row = (int)(e / K), r = (int)(e - (long long)row * K);
    const u64 k = keys[(size_t)row * N + r];
    const int col = (int)(k & 0xffffffull);
    const size_t src = (size_t)row * N + col;
    out_scores[e] = scores[src];
    if (out_boxes) out_boxes[e] = boxes[src];
    if (out_idx) out_idx[e] = col;
}

}  // namespace ronk

using namespace ronk;

extern "C" size_t ronk_sort_rows_workspace_bytes(int S, int N) {
    if (S < 1 || N < 1) return 0;
    const long long n = (long long)S * N;
    const int nblocks = (int)((n + kSortTile - 1) / kSortTile);
    return 2 * align256((size_t)n * 8) + align256((size_t)256 * nblocks * 4);
}

extern "C" int ronk_sort_rows(const float* scores, const float* boxes, int S, int N, int K, float* out_scores,
                              float* out_boxes, int32_t* out_idx, void* ws, size_t ws_bytes, void* stream) {
    RONK_REQUIRE(scores && out_scores && ws, RONK_EINVAL, "ronk_sort_rows: NULL argument");
    RONK_REQUIRE((boxes == nullptr) == (out_boxes == nullptr), RONK_EINVAL, "ronk_sort_rows: boxes and out_boxes go together");
    RONK_REQUIRE(S >= 1 && S <= 256 && N >= 1 && N < (1 << 24) && K >= 1 && K <= N, RONK_EINVAL,
                 "ronk_sort_rows: 1 <= S <= 256, 1 <= K <= N < 2^24");
    RONK_REQUIRE(!boxes || (((uintptr_t)boxes % 16) == 0 && ((uintptr_t)out_boxes % 16) == 0), RONK_EINVAL,
                 "ronk_sort_rows: box pointers must be 16-byte aligned");
    const long long n = (long long)S * N;
    RONK_REQUIRE(n < (1ll << 31) - kSortTile, RONK_ELIMIT, "ronk_sort_rows: too many elements");
    RONK_REQUIRE(ws_bytes >= ronk_sort_rows_workspace_bytes(S, N), RONK_EINVAL, "ronk_sort_rows: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int nblocks = (int)((n + kSortTile - 1) / kSortTile);
    unsigned char* w = (unsigned char*)ws;
    u64* buf[2] = {(u64*)w, (u64*)(w + align256((size_t)n * 8))};
    unsigned* hist = (unsigned*)(w + 2 * align256((size_t)n * 8));
    rows_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(scores, N, n, buf[0]);
    RONK_LAUNCHED();
    int cur = 0;
    for (int byte = 0; byte < 8; ++byte) {
        // bytes 0..2: column (those above the width of N - 1 are zero), 3..6: score, 7: row
        if (byte < 3 && ((unsigned)(N - 1) >> (8 * byte)) == 0u) continue;
        if (byte == 7 && S == 1) continue;
        if (int rc = radix_pass(buf[cur], buf[cur ^ 1], n, 8 + byte, hist, nblocks, st)) return rc;
        cur ^= 1;
    }
    const long long total = (long long)S * K;
    rows_gather_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(buf[cur], scores, (const float4*)boxes, N, K, total,
                                                                        out_scores, (float4*)out_boxes, out_idx);
    RONK_LAUNCHED();
    return RONK_OK;
}

extern "C" size_t ronk_average_precision_workspace_bytes(long long n, int C) {
    if (n < 0 || C < 2) return 0;
    return ap_layout(n, C - 1).total;
}

extern "C" int ronk_average_precision_records(const uint64_t* records, long long n, const int64_t* n_gt, int C,
                                              const double* thresholds, int n_thresholds, double* out_ap07,
                                              double* out_ap12, int32_t* out_offsets, uint64_t* out_sorted,
                                              double* out_precision, double* out_recall, void* ws, size_t ws_bytes,
                                              void* stream) {
    RONK_REQUIRE(n_gt && out_ap07 && out_ap12 && ws && thresholds, RONK_EINVAL, "ronk_average_precision_records: NULL argument");
    RONK_REQUIRE(n >= 0 && (n == 0 || records), RONK_EINVAL, "ronk_average_precision_records: NULL records");
    RONK_REQUIRE(C >= 2 && C - 1 <= 65535, RONK_EINVAL, "ronk_average_precision_records: 2 <= C <= 65536");
    RONK_REQUIRE(n < (1ll << 31) - kSortTile, RONK_ELIMIT, "ronk_average_precision_records: too many records");
    RONK_REQUIRE(n_thresholds >= 1 && n_thresholds <= kApMaxThr, RONK_EINVAL, "ronk_average_precision_records: 1..16 thresholds");
    const int CM = C - 1;
    const ApLayout L = ap_layout(n, CM);
    RONK_REQUIRE(ws_bytes >= L.total, RONK_EINVAL, "ronk_average_precision_records: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* w = (unsigned char*)ws;
    u64* buf[2] = {(u64*)(w + L.buf0), (u64*)(w + L.buf1)};
    unsigned* hist = (unsigned*)(w + L.hist);

    // ---- 1. stable LSD radix sort: 4 score bytes, then as many class bytes as C needs
    // padding entries carry class 0xffffff: every byte of it is 0xff, so they sort behind the real classes as long as
    // the largest real class (C - 2) is below 0xff.. on the bytes that are looked at
    int class_passes = 1;
    while (class_passes < 3 && (long long)CM > (1ll << (8 * class_passes)) - 1) ++class_passes;
    const u64* src = (const u64*)records;
    int cur = 0;
    if (n > 0) {
        for (int pass = 0; pass < 4 + class_passes; ++pass) {
            if (int rc = radix_pass(src, buf[cur], n, pass, hist, L.nblocks, st)) return rc;
            src = buf[cur];
            cur ^= 1;
        }
    }

    // ---- 2. per-class curves and the two AP numbers
    ApParams p;
    p.rec = src;
    p.n = n;
    p.CM = CM;
    p.T = L.T;
    p.n_gt = (const long long*)n_gt;
    p.off = (int*)(w + L.off);
    p.part = (uint2*)(w + L.part);
    p.pmax = (double*)(w + L.pmax);
    p.part12 = (double*)(w + L.part12);
    p.prec = (double*)(w + L.prec);
    p.ctp = (unsigned*)(w + L.ctp);
    p.sidx = (int*)(w + L.sidx);
    p.v07 = (double*)(w + L.v07);
    p.n_thr = n_thresholds;
    for (int i = 0; i < kApMaxThr; ++i) p.thr[i] = i < n_thresholds ? thresholds[i] : 0.0;
    p.ap07 = out_ap07;
    p.ap12 = out_ap12;
    p.out_precision = out_precision;
    p.out_recall = out_recall;
    ap_offsets_kernel<<<(CM + 1 + 127) / 128, 128, 0, st>>>(p);
    RONK_LAUNCHED();
    const dim3 grid((unsigned)L.T, (unsigned)CM);
    ap_partial_kernel<<<grid, kApThreads, 0, st>>>(p);
    RONK_LAUNCHED();
    ap_partial_scan_kernel<<<CM, kApThreads, 0, st>>>(p);
    RONK_LAUNCHED();
    ap_precision_kernel<<<grid, kApThreads, 0, st>>>(p);
    RONK_LAUNCHED();
    ap_suffix_kernel<<<CM, kApThreads, 0, st>>>(p);
    RONK_LAUNCHED();
    ap_envelope_kernel<<<grid, kApThreads, 0, st>>>(p);
    RONK_LAUNCHED();
    ap_final_kernel<<<(CM + 127) / 128, 128, 0, st>>>(p);
    RONK_LAUNCHED();
    if (out_offsets) RONK_CUDA(cudaMemcpyAsync(out_offsets, p.off, (size_t)(CM + 1) * 4, cudaMemcpyDeviceToDevice, st));
    if (out_sorted && n > 0) RONK_CUDA(cudaMemcpyAsync(out_sorted, src, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
    return RONK_OK;
}
