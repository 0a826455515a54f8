// K1 (grid form): fused all-layer GT->anchor matching + target encoding for anchors that lie on the
// regular grids the reference generates (RONNet.anchors / SSDNet.anchors).  ONE launch per batch.
//
// Reference: nets/ssd_common.py:27-47 (iou_matrix), :49-75 (do_dual_max_match), :77-147
// (tf_ssd_bboxes_encode_layer), nets/ron_vgg_320.py:686,708 (objectness label).
// Spec: SURVEY.md Appendix A.3/A.4.  Results are bit-exact against oracle/ron_oracle.py.
//
// What the grid buys.  Inside one layer the anchors of one shape `a` (a "plane") are H x W copies
// of one box translated by the layer step, and every float32 the reference derives from them is
// separable: ymin/ymax (second-trip corners, ssd_common.py:105-108) depend on (row, a) only,
// xmin/xmax on (col, a) only, the inside mask (:112-115) is a rectangle of cells per plane, and both
// corner sequences are monotone.  The anchor handle stores those row / column tables (checked
// bit for bit against the per-anchor tables when the handle is built).  Therefore
//   * the cells of a plane whose anchor intersects a GT box form a rectangle [r0,r1] x [c0,c1], found
//     by two binary searches per axis -- exactly the pairs with a non-zero overlap are enumerated
//     (64 k per image for RON-320 with 1-50 GT boxes, against 350 k anchor x GT pairs);
//   * intersection = hy(row) * wx(col) and anchor area = hA(row) * wA(col): same float32 operations,
//     same order and same roundings as areas() / intersection() on the per-anchor corners.
//
// Work item = (band of rows of one layer, image) -- a contiguous range of flat anchor indices, so the
// outputs are written coalesced -- or, for the coarse layers where a handful of anchors survive the
// border mask, (flat range, image) swept densely with lane = inside anchor.  A CTA owns the per-anchor
// state (iou bits << 32 | gt) of its item in shared memory; a WARP owns a plane (or a row sub-band of
// it) and walks the GT boxes in ascending order, so "first GT wins" (tf.argmax) is a strict '>' on
// plain shared-memory reads and writes: no atomics on the per-anchor side.  Lanes tile the rectangle
// as floor(32 / width) rows x width columns per step; wx is computed once per (GT, plane), hy once per
// step.  Per GT box the best overlap and the lowest anchor that reaches it is a packed u64
// (iou bits << 32 | ~flat anchor) max: per lane in registers, per warp by REDUX only when a lane can
// reach the box's current best, per CTA by a shared-memory atomicMax, per image by one global
// atomicMax per (item, GT) whose value improved.  The CTA that finishes an image LAST applies "the
// lowest GT index claims its best anchor" (ssd_common.py:67-75) and restores the workspace to zero.
#include <stdlib.h>

#include "encode_common.cuh"

namespace ronk {

#ifndef RONK_GRID_MINB
#define RONK_GRID_MINB 3
#endif
constexpr int kPosCap = 512;         // anchors matched by threshold listed per item (more: encoded in place)
constexpr unsigned kFull = 0xffffffffu;
constexpr int kGridMaxThreads = 320; // 10 warps: a warp per plane (or row sub-band of a plane)
constexpr int kGtRec = 32;           // bytes per staged GT box: corners (16), best key (8), area (4), pad (4)

// floor(65536 / w) + 1 for w = 1..32: (n * c_inv16[w]) >> 16 == n / w for n < 2048 (checked exhaustively)
__constant__ unsigned c_inv16[33] = {0,    65537, 32769, 21846, 16385, 13108, 10923, 9363, 8193, 7282, 6554,
                                     5958, 5462,  5042,  4682,  4370,  4097,  3856,  3641, 3450, 3277, 3121,
                                     2979, 2850,  2731,  2622,  2521,  2428,  2341,  2260, 2185, 2115, 2049};

// ---- shared memory through 32-bit addresses (one base register instead of a 64-bit pointer per array)
__device__ __forceinline__ float4 lds128(unsigned a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float2 lds64f(unsigned a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ unsigned lds32(unsigned a) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts64(unsigned a, unsigned x, unsigned y) {
    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void atoms_max64(unsigned a, u64 v) {
    asm volatile("red.shared.max.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
}

// max(min(hi_a, hi_b) - max(lo_a, lo_b), 0) in the reference's op order (ssd_common.py:38-43)
template <bool NICE>
__device__ __forceinline__ float overlap_1d(float lo_a, float hi_a, float lo_b, float hi_b) {
    const float d = fminf(hi_a, hi_b) - fmaxf(lo_a, lo_b);
    return NICE ? __saturatef(d) : fmaxf(d, 0.f);      // d <= GT side <= 1 when NICE (checked per image)
}

// inter / ((ga + aa) - inter), ssd_common.py:44-47.  The anchors of a grid handle have a positive area (checked
// when the handle is built) and inter <= aa, so the union is > 0: no `union == 0` case here.
template <bool NICE>
__device__ __forceinline__ float overlap_ratio(float inter, float ga, float aa) {
    const float uni = (ga + aa) - inter;
    if (NICE) {
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(uni));
        const float e = __fmaf_rn(-uni, r, 1.f);
        r = __fmaf_rn(r, e, r);
        const float q = __fmul_rn(inter, r);
        const float rem = __fmaf_rn(-uni, q, inter);
        return __fmaf_rn(r, rem, q);                   // == div.rn (see div_overlap_nice in common.cuh)
    }
    return div_overlap(inter, uni);
}

// Byte offsets of the arrays in the CTA's dynamic shared memory (32-bit shared addresses)
struct GridSmem {
    unsigned state;    // uint2 [S]   per anchor of the item: (x = gt index, y = iou bits)
    unsigned row;      // float4 [A * rows] (ymin, ymax, hA, -)   mode 0
    unsigned col;      // float4 [A * W]    (xmin, xmax, wA, -)   mode 0
    unsigned gt;       // kGtRec bytes per GT box: corners, best key (iou bits << 32 | ~flat anchor), area, -
};

// first index in [lo, hi) for which pred is false (pred: true ... true false ... false)
template <class F>
__device__ __forceinline__ int partition_point(int lo, int hi, F pred) {
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (pred(mid)) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// One warp, one plane `a`, rows [rs_lo, rs_hi] (band-relative, inside the plane's inside rows), inside
// columns [ci0, ci1]: walks the GT boxes in ascending order.
template <bool NICE>
__device__ __forceinline__ void rect_task(const GridSmem& s, int a, int rows, int W, int A, int pstride, unsigned n_plane0,
                                          int rs_lo, int rs_hi, int ci0, int ci1, int G) {
    const int lane = threadIdx.x & 31;
    const unsigned row = s.row + (unsigned)(a * rows) * 16u;
    const unsigned col = s.col + (unsigned)(a * W) * 16u;
    const unsigned plane = s.state + (unsigned)(a * pstride) * 8u;
#pragma unroll 1
    for (int g0 = 0; g0 < G; g0 += 32) {
        // ---- lane-parallel: rectangle of GT box g0 + lane in this plane and everything else that is the same
        // for all lanes of a unit.  u0 = r0 | r1 << 8 | c0 << 16 | c1 << 24, u1 = inv(width) (0: no rectangle)
        unsigned u0 = 0u, u1 = 0u;
        if (g0 + lane < G) {
            const float4 t = lds128(s.gt + (unsigned)(g0 + lane) * kGtRec);
            const int r0 = partition_point(rs_lo, rs_hi + 1, [&](int r) { return !(__uint_as_float(lds32(row + r * 16 + 4)) > t.x); });
            const int r1 = partition_point(rs_lo, rs_hi + 1, [&](int r) { return __uint_as_float(lds32(row + r * 16)) < t.z; }) - 1;
            const int c0 = partition_point(ci0, ci1 + 1, [&](int c) { return !(__uint_as_float(lds32(col + c * 16 + 4)) > t.y); });
            const int c1 = partition_point(ci0, ci1 + 1, [&](int c) { return __uint_as_float(lds32(col + c * 16)) < t.w; }) - 1;
            if ((r0 <= r1) && (c0 <= c1)) {
                u0 = (unsigned)r0 | ((unsigned)r1 << 8) | ((unsigned)c0 << 16) | ((unsigned)c1 << 24);
                u1 = c_inv16[min(c1 - c0 + 1, 32)];
            }
        }
        unsigned todo = __ballot_sync(kFull, u1 != 0u);
#pragma unroll 1
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const unsigned gq = (unsigned)(g0 + src);
            const unsigned rc = __shfl_sync(kFull, u0, src);
            const unsigned inv = __shfl_sync(kFull, u1, src);
            const int r0 = (int)(rc & 255u), r1 = (int)((rc >> 8) & 255u);
            const int c0 = (int)((rc >> 16) & 255u), c1 = (int)(rc >> 24);
            const unsigned rec = s.gt + gq * kGtRec;
            const float4 t = lds128(rec);
            const float4 u = lds128(rec + 16);                              // (best key lo, hi, area, -)
            const float ga = u.z;
            const unsigned cur = max(__float_as_uint(u.y), 1u);             // >= 1: zero overlaps never pass
            // lanes tile the rectangle as rpi = 32 / width rows x width columns per step (width <= 32 here; a wider
            // rectangle takes extra passes of 32 columns)
            const int rpi = (int)(inv >> 11) & 63;                          // (32 * inv) >> 16
            const int rl = (int)(((unsigned)lane * inv) >> 16);             // lane / width
            const int wd = min(c1 - c0 + 1, 32);
            const int cl = lane - rl * wd;
            const unsigned rstep = (unsigned)rpi * 16u, sstep = (unsigned)(rpi * W) * 8u;
#pragma unroll 1
            for (int cc = c0; cc <= c1; cc += 32) {
                const int c = min(cc + cl, c1);
                const float4 ct = lds128(col + c * 16);
                const float wx = overlap_1d<NICE>(t.y, t.w, ct.x, ct.y);
                const float wA = ct.z;
                unsigned mymax = 0u;
                int myr = 0;
                // lane-divergent trip count: lanes past the last row fall out early, spare lanes never enter
                int r = (rl >= rpi || cc + cl > c1) ? r1 + 1 : r0 + rl;
                unsigned rp = row + r * 16, sp = plane + (unsigned)(r * W + c) * 8u;
                for (; r <= r1; r += rpi, rp += rstep, sp += sstep) {
                    const float4 rt = lds128(rp);
                    const float hy = overlap_1d<NICE>(t.x, t.z, rt.x, rt.y);
                    const float inter = hy * wx;
                    const float iou = overlap_ratio<NICE>(inter, ga, wA * rt.z);
                    const unsigned bits = __float_as_uint(iou);
                    if (bits > lds32(sp + 4)) sts64(sp, gq, bits);                // ascending gq: first GT wins ties
                    if (bits > mymax) { mymax = bits; myr = r; }                 // ascending r: lowest anchor wins ties
                }
                // per-GT (max, lowest anchor).  IoU >= 0, so float bits order as integers.
                if (__any_sync(kFull, mymax >= cur)) {
                    const unsigned m = __reduce_max_sync(kFull, mymax);
                    const unsigned cand = (mymax == m) ? n_plane0 + (unsigned)((myr * W + c) * A) : 0xffffffffu;
                    const unsigned first = __reduce_min_sync(kFull, cand);
                    if (lane == 0) atoms_max64(rec + 16, ((u64)m << 32) | (u64)(0xffffffffu - first));
                }
                __syncwarp();
            }
        }
    }
}

// One warp, 32 consecutive inside anchors (compact order) of a dense item against every GT box.
template <bool NICE>
__device__ __forceinline__ void dense_task(const EncodeParams& p, const GridSmem& s, int c_first, int c_hi, int n_lo, int G) {
    const int lane = threadIdx.x & 31;
    const int c = c_first + lane;
    const bool in = c < c_hi;
    const float4 a = in ? p.ccor[c] : make_float4(CUDART_INF_F, CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
    const unsigned n = in ? (unsigned)p.inside_idx[c] : 0xffffffffu;
    const float area = in ? (a.w - a.y) * (a.z - a.x) : 0.f;
    float best = 0.f;
    unsigned bestg = 0;
    for (int g = 0; g < G; ++g) {
        const unsigned rec = s.gt + (unsigned)g * kGtRec;
        const float4 t = lds128(rec);
        const float4 u = lds128(rec + 16);                                  // (best key lo, hi, area, -)
        const float ga = u.z;
        const unsigned cur = max(__float_as_uint(u.y), 1u);
        const float iou = iou_ref<NICE>(t, ga, a, area);
        const unsigned bits = __float_as_uint(iou);
        if (iou > best) { best = iou; bestg = (unsigned)g; }
        if (__any_sync(kFull, bits >= cur)) {
            const unsigned m = __reduce_max_sync(kFull, bits);
            const unsigned first = __reduce_min_sync(kFull, bits == m ? n : 0xffffffffu);
            if (lane == 0) atoms_max64(rec + 16, ((u64)m << 32) | (u64)(0xffffffffu - first));
        }
    }
    if (in) sts64(s.state + (n - (unsigned)n_lo) * 8u, bestg, __float_as_uint(best));
}

// (not inlined: the sweep gets its own register allocation, away from the output and forcing code)
template <bool NICE>
__device__ __noinline__ void run_tasks(const EncodeParams& p, const GridItem& it, const GridSmem& s, int G) {
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    if (it.mode == 0) {
        const int A = it.A, nsub = it.nsub, rows = it.rows, W = it.W, r_lo = it.r_lo, pstride = it.pstride;
        const int pl_base = it.pl_base;
        const unsigned n_band0 = (unsigned)(it.layer_n0 + r_lo * W * A);
        const int ntasks = A * nsub;
#pragma unroll 1
        for (int t = warp; t < ntasks; t += nwarps) {
            const int a = t / nsub, sub = t - a * nsub;
            const int4 pl = __ldg(p.planes + pl_base + a);               // inside rows [x, y], columns [z, w] of the plane
            // rows of this sub-band, clipped to the plane's inside rows, relative to the band
            const int b_lo = r_lo + (sub * rows) / nsub, b_hi = r_lo + ((sub + 1) * rows) / nsub - 1;
            const int rs_lo = max(b_lo, pl.x) - r_lo, rs_hi = min(b_hi, pl.y) - r_lo;
            if (rs_lo > rs_hi || pl.z > pl.w) continue;
            rect_task<NICE>(s, a, rows, W, A, pstride, n_band0 + (unsigned)a, rs_lo, rs_hi, pl.z, pl.w, G);
        }
    } else {
        const int c_lo = it.c_lo, c_hi = it.c_hi, n_lo = it.n_lo;
        const int ntasks = (c_hi - c_lo + 31) >> 5;
#pragma unroll 1
        for (int t = warp; t < ntasks; t += nwarps) dense_task<NICE>(p, s, c_lo + 32 * t, c_hi, n_lo, G);
    }
}

// the rare case of more than kPosCap threshold matches in one item: encoded where it is found (kept out of line)
__device__ __noinline__ void encode_in_place(const EncodeParams& p, float4 gb, int n, float4* dst) {
    *dst = encode_loc(gb, p.enc[n], p);
}

__device__ __noinline__ void force_image_call(const EncodeParams& p, int b, int G, int* s_mem, int nt) {
    force_image(p, b, G, s_mem, s_mem + p.gcap, nt);
}

#ifdef RONK_ENC_TRACE
// profiling build only (tools/enc_trace.py): per CTA 8 words {start ns, SM id, end ns, after staging, after the
// sweep, after the output stores, after the tile counter, unused}
__device__ unsigned long long* g_grid_trace;
__device__ __forceinline__ unsigned long long trace_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define RONK_TRACE_MARK(k)                                                                          \
    do {                                                                                            \
        if (g_grid_trace && threadIdx.x == 0)                                                       \
            g_grid_trace[8 * ((size_t)blockIdx.y * gridDim.x + blockIdx.x) + (k)] = trace_now();    \
    } while (0)
#else
#define RONK_TRACE_MARK(k) do { } while (0)
#endif

// order[k] = the image with the k-th largest GT count (ties: lower index first)
__global__ void __launch_bounds__(256)
rank_images_kernel(const int* __restrict__ counts, int B, int Gmax, int* __restrict__ order) {
    __shared__ int s_c[256];
    const int b = blockIdx.x * 256 + threadIdx.x;
    int g = b < B ? counts[b] : 0;
    g = g < 0 ? 0 : (g > Gmax ? Gmax : g);
    int rank = 0;
    for (int b0 = 0; b0 < B; b0 += 256) {
        int v = b0 + threadIdx.x < B ? counts[b0 + threadIdx.x] : -1;
        s_c[threadIdx.x] = v < 0 ? (b0 + threadIdx.x < B ? 0 : -1) : (v > Gmax ? Gmax : v);
        __syncthreads();
        const int lim = min(256, B - b0);
        for (int j = 0; j < lim; ++j) {
            const int v2 = s_c[j];
            rank += (v2 > g || (v2 == g && b0 + j < b)) ? 1 : 0;
        }
        __syncthreads();
    }
    if (b < B) order[rank] = b;
}

// grid (items, B): blockIdx.x walks the handle's item table (heavy items first), blockIdx.y the images by weight
__global__ void __launch_bounds__(kGridMaxThreads, RONK_GRID_MINB)
match_encode_grid_kernel(const __grid_constant__ EncodeParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ GridItem s_item;
    __shared__ unsigned s_npos;
    __shared__ int s_last;

    const int tid = threadIdx.x, nt = blockDim.x;
    // the grid is image-major over the images sorted by descending GT count: the CTAs still to start when the
    // machine drains are the cheapest ones
    const int b = __ldg(p.order + blockIdx.y);
#ifdef RONK_ENC_TRACE
    if (g_grid_trace && tid == 0) {
        unsigned sm;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        g_grid_trace[8 * ((size_t)blockIdx.y * gridDim.x + blockIdx.x) + 1] = sm;
    }
    RONK_TRACE_MARK(0);
#endif
    if (tid < (int)(sizeof(GridItem) / 4))
        reinterpret_cast<int*>(&s_item)[tid] = __ldg(reinterpret_cast<const int*>(p.gitems + blockIdx.x) + tid);
    if (tid == 0) s_npos = 0u;
    // loads that depend on nothing else go first: the image's GT slot of this thread, the image-wide bests
    const float4* gtb = p.gt_boxes + (size_t)b * p.Gmax;
    const long long* gtl = p.gt_labels + (size_t)b * p.Gmax;
    const u64* wsk = p.ws_keys + (size_t)b * p.Gmax;
    const bool spec = tid < p.Gmax;
    const float4 v_spec = spec ? gtb[tid] : make_float4(0.f, 0.f, 0.f, 0.f);
    const long long l_spec = spec ? gtl[tid] : 0ll;
    const u64 k_spec = spec ? __ldcg(wsk + tid) : 0ull;
    int G = p.gt_counts[b];
    G = G < 0 ? 0 : (G > p.Gmax ? p.Gmax : G);
    __syncthreads();
    const GridItem& it = s_item;
    const int mode = it.mode, n_lo = it.n_lo, n_hi = it.n_hi;

    // dynamic shared memory: state | row table | column table | GT records | start keys | labels | positives
    const int S = mode == 0 ? it.A * it.pstride : n_hi - n_lo;
    const int S2 = (S + 1) & ~1;                                   // keeps what follows 16-byte aligned
    const unsigned o_row = (unsigned)S2 * 8u;
    const unsigned o_col = o_row + (mode == 0 ? (unsigned)(it.A * it.rows) * 16u : 0u);
    const unsigned o_gt = o_col + (mode == 0 ? (unsigned)(it.A * it.W) * 16u : 0u);
    const unsigned o_ginit = o_gt + (unsigned)p.gcap * kGtRec;
    const unsigned o_lab = o_ginit + (unsigned)p.gcap * 8u;
    const unsigned o_pos = o_lab + (unsigned)p.gcap * 8u;
    GridSmem s;
    s.state = (unsigned)__cvta_generic_to_shared(smem);
    s.row = s.state + o_row;
    s.col = s.state + o_col;
    s.gt = s.state + o_gt;
    u64* s_ginit = reinterpret_cast<u64*>(smem + o_ginit);
    long long* s_lab = reinterpret_cast<long long*>(smem + o_lab);
    unsigned* s_pos = reinterpret_cast<unsigned*>(smem + o_pos);

    // ---- prologue: zero the state, copy the item's row / column tables, stage the GT boxes
    {
        uint4* z = reinterpret_cast<uint4*>(smem);
        #pragma unroll 1
        for (int i = tid; i < S2 / 2; i += nt) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (mode == 0) {
        const int rows = it.rows, nr = it.A * rows;
        float4* s_row = reinterpret_cast<float4*>(smem + o_row);
        const float4* src = p.rowtab + it.rt_base + it.r_lo;
        #pragma unroll 1
        for (int i = tid; i < nr; i += nt) {
            const int a = i / rows, r = i - a * rows;
            s_row[i] = __ldg(src + a * it.H + r);
        }
        const int nc = it.A * it.W;
        float4* s_col = reinterpret_cast<float4*>(smem + o_col);
        const float4* csrc = p.coltab + it.ct_base;
        #pragma unroll 1
        for (int i = tid; i < nc; i += nt) s_col[i] = __ldg(csrc + i);
    }
    bool nice = true;
    #pragma unroll 1
    for (int g = tid; g < G; g += nt) {
        const float4 v = (g == tid) ? v_spec : gtb[g];
        // start from the image-wide best published so far (any stale value is a valid lower bound):
        // only overlaps that can still win reach the reduction path
        const u64 k0 = (g == tid) ? k_spec : __ldcg(wsk + g);
        unsigned char* rec = smem + o_gt + (size_t)g * kGtRec;
        *reinterpret_cast<float4*>(rec) = v;
        *reinterpret_cast<u64*>(rec + 16) = k0;
        *reinterpret_cast<float2*>(rec + 24) = make_float2((v.w - v.y) * (v.z - v.x), 0.f);
        s_ginit[g] = k0;
        s_lab[g] = (g == tid) ? l_spec : gtl[g];
        nice = nice && nice_coord(v.x) && nice_coord(v.y) && nice_coord(v.z) && nice_coord(v.w) &&
               (v.z - v.x) <= 1.f && (v.w - v.y) <= 1.f;
    }
    nice = __syncthreads_and(nice && p.anchors_nice) != 0;
    RONK_TRACE_MARK(3);

    // ---- sweep
    if (nice) run_tasks<true>(p, it, s, G); else run_tasks<false>(p, it, s, G);
    RONK_TRACE_MARK(7);          // thread 0's warp done with its own tasks
    __syncthreads();
    RONK_TRACE_MARK(4);

    // ---- publish the per-GT bests this item improved
    #pragma unroll 1
    for (int g = tid; g < G; g += nt) {
        const u64 v = *reinterpret_cast<const u64*>(smem + o_gt + (size_t)g * kGtRec + 16);
        if (v > s_ginit[g]) atomicMax(p.ws_keys + (size_t)b * p.Gmax + g, v);
    }

    // ---- label + store the item's flat anchor range (forced anchors are rewritten by the CTA that finishes
    // the image last).  Anchors matched by threshold are listed; their localisations (4 IEEE divisions and 2
    // double-precision logs each) are encoded afterwards, one listed anchor per thread.
    {
        const unsigned A = (unsigned)it.A, pstride = (unsigned)it.pstride;
        const unsigned magicA = mode == 0 ? 0xffffffffu / A + 1u : 0u;   // n / A for n < 2^16
        const float low = p.low, high = p.high;
        const bool ib = p.ignore_between != 0;
        const size_t o0 = (size_t)b * p.N + n_lo;
        long long* o_labels = p.out_labels + o0;
        float4* o_loc = p.out_loc + o0;
        float* o_scores = p.out_scores + o0;
        int* o_matched = p.out_matched ? p.out_matched + o0 : nullptr;
        int* o_obj = p.out_obj ? p.out_obj + o0 : nullptr;
        const unsigned cnt = (unsigned)(n_hi - n_lo);
        #pragma unroll 1
        for (unsigned local = tid; local < cnt; local += nt) {
            unsigned sidx = local;
            if (mode == 0) {
                const unsigned cell = __umulhi(local, magicA);
                sidx = (local - cell * A) * pstride + cell;
            }
            const unsigned sa = s.state + sidx * 8u;
            const unsigned a2g = lds32(sa);
            const float mv = __uint_as_float(lds32(sa + 4));
            const bool less = mv < low;
            const bool between = (mv < high) && !less;
            const bool neg = ib ? less : between;
            const bool ign = ib ? between : less;
            int mi = ign ? -2 : (neg ? -1 : (int)a2g);
            if (G == 0) mi = -1;
            long long label = mi < -1 ? -1ll : 0ll;
            if (mi >= 0) {
                label = s_lab[a2g];
                const unsigned slot = atomicAdd(&s_npos, 1u);
                if (slot < (unsigned)kPosCap) s_pos[slot] = (local << 10) | a2g;
                else encode_in_place(p, *reinterpret_cast<const float4*>(smem + o_gt + (size_t)a2g * kGtRec), n_lo + (int)local, o_loc + local);
                if (!p.gt_max_first) atomicOr(p.ws_claimed + (size_t)b * p.Gmax + a2g, 1u);
            } else {
                o_loc[local] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            o_labels[local] = label;
            o_scores[local] = mv;
            if (o_matched) o_matched[local] = mi;
            if (o_obj) o_obj[local] = label > 0 ? 1 : 0;
        }
        __syncthreads();
        const unsigned npos = min(s_npos, (unsigned)kPosCap);
        #pragma unroll 1
        for (unsigned k = tid; k < npos; k += nt) {
            const unsigned e = s_pos[k];
            const unsigned local = e >> 10;
            o_loc[local] = encode_loc(*reinterpret_cast<const float4*>(smem + o_gt + (size_t)(e & 1023u) * kGtRec),
                                      p.enc[n_lo + local], p);
        }
    }

    // ---- the last item of the image applies the per-GT forcing.  Barrier first, then one thread fences and
    // bumps the counter (release pattern of a grid-wide barrier).
    __syncthreads();
    RONK_TRACE_MARK(5);
    if (tid == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(p.ws_count + b, 1u);
        s_last = (prev == (unsigned)p.tiles - 1u) ? 1 : 0;
    }
    __syncthreads();
    RONK_TRACE_MARK(6);
    if (s_last) {
        __threadfence();
        // the state array is free now: reuse its head as the two int arrays of the forcing step
        force_image_call(p, b, G, reinterpret_cast<int*>(smem), nt);
    }
#ifdef RONK_ENC_TRACE
    __syncthreads();
    RONK_TRACE_MARK(2);
#endif
}

#ifdef RONK_ENC_TRACE
}  // namespace ronk
extern "C" int ronk_debug_set_enc_trace_grid(void* buf) {
    RONK_CUDA(cudaMemcpyToSymbol(ronk::g_grid_trace, &buf, sizeof(buf)));
    return RONK_OK;
}
namespace ronk {
#endif

int launch_rank_images(const EncodeParams& p, int B, cudaStream_t st) {
    rank_images_kernel<<<(B + 255) / 256, 256, 0, st>>>(p.gt_counts, B, p.Gmax, const_cast<int*>(p.order));
    RONK_LAUNCHED();
    return RONK_OK;
}

int launch_match_encode_grid(const ronk_anchors* h, EncodeParams& p, int B, cudaStream_t st) {
    // Item table: bands of ~4096 anchors measured best at every batch size this kernel is dispatched for (finer cuts
    // pay the per-CTA fixed cost -- prologue, fence + counter -- more often); the other two tables stay selectable.
    int table = 0;
    if (const char* e = getenv("RONK_ENC_TABLE")) {             // tuning knob
        const int v = atoi(e);
        if (v >= 0 && v <= 2) table = v;
    }
    p.gitems = (const GridItem*)h->d_gitems[table];
    p.tiles = h->n_gitems[table];
    p.rowtab = (const float4*)h->d_rowtab;
    p.coltab = (const float4*)h->d_coltab;
    p.planes = (const int4*)h->d_planes;
    p.key_flat = 1;
    RONK_REQUIRE(B <= 65535, RONK_ELIMIT, "ronk_match_encode: more than 65535 images in one call");
    if (int rc = launch_rank_images(p, B, st)) return rc;
    const size_t smem = h->gitems_smem[table] + (size_t)p.gcap * (kGtRec + 8 + 8) + (size_t)kPosCap * 4;
    RONK_REQUIRE(smem <= 200 * 1024, RONK_ELIMIT, "ronk_match_encode: item state does not fit in shared memory");
    if (smem > 48 * 1024)
        RONK_CUDA(cudaFuncSetAttribute(match_encode_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    match_encode_grid_kernel<<<dim3((unsigned)p.tiles, (unsigned)B), h->grid_threads, smem, st>>>(p);
    RONK_LAUNCHED();
    return RONK_OK;
}

}  // namespace ronk
