// K3: batched greedy NMS, one warp per (image, class) segment.
//
// Reference: tf_extended/bboxes.py:173-234 (bboxes_nms) and :262-302 (bboxes_nms_batch).
// Spec: SURVEY.md Appendix A.7.  Bit-exact (kept set and order) against oracle/ron_oracle.py.
//
// The reference's loop "pick the first live box, kill everything it overlaps, stop after M
// picks" is equivalent to: walk the score-sorted candidates once and keep a candidate iff no
// previously KEPT box suppresses it and fewer than M are kept.  The warp walks the candidates
// in chunks of 32 (one per lane): every lane tests its candidate against the kept list in
// shared memory (broadcast reads), a 32x32 suppression bit-matrix is built inside the chunk,
// and the chunk is resolved with ballots/shuffles.  Memory is O(M) per segment for any K, the
// walk stops as soon as M boxes are kept.  overlap = inter / min(area_j, area_i) ('min',
// the reference default) or inter / ((area_j - inter) + area_i) ('union'), safe_divide
// (0 when the denominator is <= 0), keep test is strict: overlap < threshold.
#include <stdlib.h>
#include "common.cuh"
#include "topk.cuh"

namespace ronk {

constexpr int kNmsWarps = 4;
constexpr int kNmsPad = 8;            // kept boxes are tested in unrolled groups of this many

__host__ __device__ constexpr int nms_smem_per_warp(int M) {
    return ((M + kNmsPad + 32) * 20 + M * 4 + 15) & ~15;
}

struct NmsParams {
    const float* scores;
    const float4* boxes;
    const int* order;   // [S,K] sorted positions, or NULL when rows are already sorted
    int S, K, M, mode;
    float thr;
    float* out_scores;
    float4* out_boxes;
    int* out_idx;
    const int* only_flagged;   // [S] or NULL: segments whose flag is 0 are left alone
    int* out_short;            // [S] or NULL: 1 when all K candidates were walked and fewer than M were kept
};

// does kept box i suppress candidate j?  (j is the later one: tf_extended/bboxes.py:195-211)
// overlap = inner / den rounded to float32, suppressed iff !(overlap < thr).  The division is
// only executed when inner is within 1e-6 (relative) of thr * den: outside that band the
// rounded quotient cannot land on the other side of thr (float32 rounding moves it by at most
// 6e-8 relative and is monotonic), so the sign of inner - thr * den decides.
__device__ __forceinline__ bool suppresses(float4 bj, float vj, float4 bi, float vi, int mode, float thr,
                                           bool zero_suppresses) {
    float h = fminf(bj.z, bi.z) - fmaxf(bj.x, bi.x);
    float w = fminf(bj.w, bi.w) - fmaxf(bj.y, bi.y);
    if (h > 0.f && w > 0.f) {
        float inner = h * w;
        float den = (mode == RONK_NMS_UNION) ? ((vj - inner) + vi) : fminf(vj, vi);
        if (!(den > 0.f)) return zero_suppresses;          // safe_divide: overlap is 0
        float t = thr * den;
        float d = inner - t;
        if (fabsf(d) > t * 1e-6f) return d > 0.f;
        return !(inner / den < thr);
    }
    return zero_suppresses;   // overlap is exactly 0: suppressed only when !(0 < thr)
}

__global__ void __launch_bounds__(kNmsWarps * 32)
nms_kernel(const __grid_constant__ NmsParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int seg = blockIdx.x * kNmsWarps + warp;
    if (seg >= p.S) return;
    if (p.only_flagged && !p.only_flagged[seg]) return;
    // per warp: float4 + float for M kept (+ kNmsPad sentinels) and 32 chunk entries, + thr * vol
    const int per_warp = nms_smem_per_warp(p.M);
    unsigned char* base = smem + (size_t)warp * per_warp;
    float4* s_kbox = reinterpret_cast<float4*>(base);
    float4* s_cbox = s_kbox + p.M + kNmsPad;
    // (s_kvol sits between s_cvol and s_kt: the compiler reads it four entries at a time, up to three past `count`)
    float* s_cvol = reinterpret_cast<float*>(s_cbox + 32);
    float* s_kvol = s_cvol + 32;
    float* s_kt = s_kvol + p.M;           // thr * vol of the kept boxes (+inf for empty boxes: they suppress nothing)
    // Sentinels behind the kept list: a box no candidate intersects, with threshold +inf.  The fast loop below
    // then always runs whole groups of kNmsPad kept boxes, fully unrolled, without a bound inside the group.
    for (int i = lane; i < p.M + kNmsPad; i += 32) {
        s_kbox[i] = make_float4(2.f, 2.f, -1.f, -1.f);
        s_kt[i] = __int_as_float(0x7f800000);
    }
    __syncwarp();
    // fast 'min' loop: valid while thr > 0 and every box seen so far has all corners in [0, 1]
    bool unit = (p.mode == RONK_NMS_MIN) && (p.thr > 0.f);

    const bool zero_supp = !(0.f < p.thr);
    const size_t in0 = (size_t)seg * p.K, out0 = (size_t)seg * p.M;
    int count = 0;
    for (int c0 = 0; c0 < p.K && count < p.M; c0 += 32) {
        const int j = c0 + lane;
        const bool valid = j < p.K;
        int pos = -1;
        float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
        float score = 0.f;
        if (valid) {
            pos = p.order ? p.order[in0 + j] : j;
            box = p.boxes[in0 + pos];
            score = p.scores[in0 + pos];
        }
        const float vol = (box.w - box.y) * (box.z - box.x);
        // 1. against everything kept by earlier chunks (broadcast reads).  Branch-free fast test: the
        //    sign of inner - thr * den decides unless it is within 1e-6 relative of zero (then, or for
        //    NaNs, the chunk is redone with the exact division below).
        bool dead = !valid;
        bool unclear = false;
        unit = unit && __all_sync(full, box.x >= 0.f && box.y >= 0.f && box.z <= 1.f && box.w <= 1.f &&
                                            box.x <= 1.f && box.y <= 1.f && box.z >= 0.f && box.w >= 0.f);
        if (unit) {
            // overlap = inner / min(vol_j, vol_i) >= thr  <=>  inner >= min(thr vol_j, thr vol_i) (rounding is
            // monotonic).  Sides are <= 1, so max(., 0) is a saturating subtract (FMA pipe); the band in
            // which the exact division must decide is taken against thr vol_j >= min(.): conservative.
            const float tj = p.thr * vol;
            const float tolj = tj * 1e-6f;
            for (int i0 = 0; i0 < count; i0 += kNmsPad) {
#pragma unroll
                for (int i = 0; i < kNmsPad; ++i) {                 // entries past `count` are sentinels: d = -tj
                    const float4 kb = s_kbox[i0 + i];
                    const float kt = s_kt[i0 + i];
                    const float h = __saturatef(fminf(box.z, kb.z) - fmaxf(box.x, kb.x));
                    const float w = __saturatef(fminf(box.w, kb.w) - fmaxf(box.y, kb.y));
                    const float d = h * w - fminf(tj, kt);
                    dead |= d > 0.f;
                    unclear |= fabsf(d) <= tolj;
                }
                if (__all_sync(full, dead)) break;
            }
            if (valid && !(vol > 0.f)) { dead = false; unclear = false; }     // empty candidate: overlap is 0 everywhere
        } else {
            for (int i0 = 0; i0 < count; i0 += 8) {
                const int lim = min(8, count - i0);
#pragma unroll 4
                for (int i = 0; i < lim; ++i) {
                    const float4 kb = s_kbox[i0 + i];
                    const float kv = s_kvol[i0 + i];
                    const float h = fmaxf(fminf(box.z, kb.z) - fmaxf(box.x, kb.x), 0.f);
                    const float w = fmaxf(fminf(box.w, kb.w) - fmaxf(box.y, kb.y), 0.f);
                    const float inner = h * w;
                    const float den = (p.mode == RONK_NMS_UNION) ? ((vol - inner) + kv) : fminf(vol, kv);
                    const float t = p.thr * den;
                    const float d = inner - t;
                    const bool pos = den > 0.f;                       // else safe_divide gives overlap 0
                    dead |= pos ? (d > 0.f) : zero_supp;
                    unclear |= pos && !(fabsf(d) > t * 1e-6f);
                }
                if (__all_sync(full, dead)) break;
            }
        }
        if (__any_sync(full, unclear && valid)) {
            dead = !valid;
            for (int i = 0; i < count; ++i)
                dead |= suppresses(box, vol, s_kbox[i], s_kvol[i], p.mode, p.thr, zero_supp);
        }
        // 2. inside the chunk: walk the survivors in order; each one that is still alive is kept
        //    and kills the later lanes it overlaps
        s_cbox[lane] = box;
        s_cvol[lane] = vol;
        __syncwarp();
        unsigned alive_mask = __ballot_sync(full, !dead);
        unsigned kept = 0u;
        int room = p.M - count;
        while (alive_mask && room > 0) {
            const int i = __ffs(alive_mask) - 1;
            kept |= 1u << i;
            --room;
            float4 kb = s_cbox[i];
            float kv = s_cvol[i];
            if (lane > i && !dead) dead = suppresses(box, vol, kb, kv, p.mode, p.thr, zero_supp);
            alive_mask = __ballot_sync(full, !dead) & ~((2u << i) - 1u);
        }
        if ((kept >> lane) & 1u) {
            int r = count + __popc(kept & ((1u << lane) - 1u));
            s_kbox[r] = box;
            s_kvol[r] = vol;
            s_kt[r] = (vol > 0.f) ? p.thr * vol : __int_as_float(0x7f800000);
            p.out_scores[out0 + r] = score;
            p.out_boxes[out0 + r] = box;
            if (p.out_idx) p.out_idx[out0 + r] = pos;
        }
        count += __popc(kept);
        __syncwarp();
    }
    if (p.out_short && lane == 0) p.out_short[seg] = count < p.M ? 1 : 0;
    for (int r = count + lane; r < p.M; r += 32) {   // pad_axis (tensors.py:59-86)
        p.out_scores[out0 + r] = 0.f;
        p.out_boxes[out0 + r] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.out_idx) p.out_idx[out0 + r] = -1;
    }
}

// Few segments with long candidate lists (crowded scenes: S ~ 1 000, K in the thousands): one warp per segment leaves
// the machine at ~10 % occupancy and every chunk walks the whole kept list on one warp.  Here a CTA of W warps owns a
// segment: all warps hold the same chunk of 32 candidates (lane = candidate), warp w tests it against every W-th group of
// 8 kept boxes, the per-warp verdicts are OR-ed through shared memory, and warp 0 resolves the chunk and appends to the
// (shared) kept list exactly as nms_kernel does.  Same results: which kept box suppresses a candidate does not matter.
constexpr int kWidePad = 2 * kNmsPad;      // nms_wide_kernel tests two groups per warp and super-step
__host__ __device__ constexpr int nms_wide_smem(int M) { return ((M + kWidePad + 32) * 20 + M * 4 + 15) & ~15; }

template <int W>
__global__ void __launch_bounds__(W * 32)
nms_wide_kernel(const __grid_constant__ NmsParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned s_dead[W], s_unc[W];
    __shared__ unsigned s_step[2][W];
    __shared__ int s_count;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int seg = blockIdx.x;
    if (p.only_flagged && !p.only_flagged[seg]) return;
    float4* s_kbox = reinterpret_cast<float4*>(smem);
    float4* s_cbox = s_kbox + p.M + kWidePad;
    float* s_cvol = reinterpret_cast<float*>(s_cbox + 32);
    float* s_kvol = s_cvol + 32;
    float* s_kt = s_kvol + p.M;
    for (int i = threadIdx.x; i < p.M + kWidePad; i += W * 32) {
        s_kbox[i] = make_float4(2.f, 2.f, -1.f, -1.f);
        s_kt[i] = __int_as_float(0x7f800000);
    }
    __syncthreads();
    bool unit = (p.mode == RONK_NMS_MIN) && (p.thr > 0.f);
    const bool zero_supp = !(0.f < p.thr);
    const size_t in0 = (size_t)seg * p.K, out0 = (size_t)seg * p.M;
    int count = 0;
    for (int c0 = 0; c0 < p.K && count < p.M; c0 += 32) {
        const int j = c0 + lane;
        const bool valid = j < p.K;
        int pos = -1;
        float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
        float score = 0.f;
        if (valid) {
            pos = p.order ? p.order[in0 + j] : j;
            box = p.boxes[in0 + pos];
            score = p.scores[in0 + pos];
        }
        const float vol = (box.w - box.y) * (box.z - box.x);
        bool dead = !valid;
        bool unclear = false;
        unit = unit && __all_sync(full, box.x >= 0.f && box.y >= 0.f && box.z <= 1.f && box.w <= 1.f &&
                                            box.x <= 1.f && box.y <= 1.f && box.z >= 0.f && box.w >= 0.f);
        if (unit) {
            // super-steps of W x 2 groups of 8 kept boxes: warp w tests groups 2w, 2w + 1 of the super-step, then the
            // verdicts so far are OR-ed over the warps -- the walk ends as soon as the whole chunk is dead, as it does in
            // the one-warp kernel, instead of every warp finishing its own slice
            const float tj = p.thr * vol;
            const float tolj = tj * 1e-6f;
            int par = 0;
            for (int s0 = 0; s0 < count; s0 += W * 2 * kNmsPad) {
                const int i0 = s0 + warp * 2 * kNmsPad;
                if (i0 < count) {
#pragma unroll
                    for (int i = 0; i < 2 * kNmsPad; ++i) {             // entries past `count` are sentinels (list padded by 2 groups)
                        const float4 kb = s_kbox[i0 + i];
                        const float kt = s_kt[i0 + i];
                        const float h = __saturatef(fminf(box.z, kb.z) - fmaxf(box.x, kb.x));
                        const float w = __saturatef(fminf(box.w, kb.w) - fmaxf(box.y, kb.y));
                        const float d = h * w - fminf(tj, kt);
                        dead |= d > 0.f;
                        unclear |= fabsf(d) <= tolj;
                    }
                }
                if (s0 + W * 2 * kNmsPad >= count) break;               // last super-step: the ballots below collect it
                const unsigned dmi = __ballot_sync(full, dead);
                if (lane == 0) s_step[par][warp] = dmi;
                __syncthreads();
                unsigned dall = 0u;
#pragma unroll
                for (int w = 0; w < W; ++w) dall |= s_step[par][w];
                par ^= 1;
                if (dall == full) { dead = true; break; }               // block-uniform
            }
            if (valid && !(vol > 0.f)) { dead = false; unclear = false; }
        } else {
            for (int i0 = warp * 8; i0 < count; i0 += W * 8) {
                const int lim = min(8, count - i0);
#pragma unroll 4
                for (int i = 0; i < lim; ++i) {
                    const float4 kb = s_kbox[i0 + i];
                    const float kv = s_kvol[i0 + i];
                    const float h = fmaxf(fminf(box.z, kb.z) - fmaxf(box.x, kb.x), 0.f);
                    const float w = fmaxf(fminf(box.w, kb.w) - fmaxf(box.y, kb.y), 0.f);
                    const float inner = h * w;
                    const float den = (p.mode == RONK_NMS_UNION) ? ((vol - inner) + kv) : fminf(vol, kv);
                    const float t = p.thr * den;
                    const float d = inner - t;
                    const bool posd = den > 0.f;
                    dead |= posd ? (d > 0.f) : zero_supp;
                    unclear |= posd && !(fabsf(d) > t * 1e-6f);
                }
                if (__all_sync(full, dead)) break;
            }
        }
        const unsigned dm = __ballot_sync(full, dead), um = __ballot_sync(full, unclear && valid);
        if (lane == 0) { s_dead[warp] = dm; s_unc[warp] = um; }
        __syncthreads();
        if (warp == 0) {
            unsigned dall = 0u, uall = 0u;
#pragma unroll
            for (int w = 0; w < W; ++w) { dall |= s_dead[w]; uall |= s_unc[w]; }
            dead = (dall >> lane) & 1u;
            if (unit && valid && !(vol > 0.f)) dead = false;          // an empty candidate overlaps nothing (as above)
            if (uall) {
                dead = !valid;
                for (int i = 0; i < count; ++i)
                    dead |= suppresses(box, vol, s_kbox[i], s_kvol[i], p.mode, p.thr, zero_supp);
            }
            s_cbox[lane] = box;
            s_cvol[lane] = vol;
            __syncwarp();
            unsigned alive_mask = __ballot_sync(full, !dead);
            unsigned kept = 0u;
            int room = p.M - count;
            while (alive_mask && room > 0) {
                const int i = __ffs(alive_mask) - 1;
                kept |= 1u << i;
                --room;
                const float4 kb = s_cbox[i];
                const float kv = s_cvol[i];
                if (lane > i && !dead) dead = suppresses(box, vol, kb, kv, p.mode, p.thr, zero_supp);
                alive_mask = __ballot_sync(full, !dead) & ~((2u << i) - 1u);
            }
            if ((kept >> lane) & 1u) {
                const int r = count + __popc(kept & ((1u << lane) - 1u));
                s_kbox[r] = box;
                s_kvol[r] = vol;
                s_kt[r] = (vol > 0.f) ? p.thr * vol : __int_as_float(0x7f800000);
                p.out_scores[out0 + r] = score;
                p.out_boxes[out0 + r] = box;
                if (p.out_idx) p.out_idx[out0 + r] = pos;
            }
            if (lane == 0) s_count = count + __popc(kept);
        }
        __syncthreads();
        count = s_count;
    }
    if (p.out_short && threadIdx.x == 0) p.out_short[seg] = count < p.M ? 1 : 0;
    for (int r = count + (int)threadIdx.x; r < p.M; r += W * 32) {
        p.out_scores[out0 + r] = 0.f;
        p.out_boxes[out0 + r] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.out_idx) p.out_idx[out0 + r] = -1;
    }
}

// Long lists in which nearly every candidate is suppressed (crowded scenes: 10 000 candidates, 100-200 kept).  Measured
// on that workload: a suppressed candidate's FIRST suppressor sits at index 6 (median) / 13 (mean) of the kept list and
// within the first 32 entries for ~90 % of them, but a chunk of 32 candidates walks the list until its LAST lane is dead
// (~90 of ~140 entries).  Two stages remove most of that:
//   A. once 32 boxes are kept those 32 never change (the list is append-only): every warp takes its own chunk of 32
//      candidates and tests it against that fixed prefix only; the survivors (~10 %) are compacted, in candidate order,
//      into a queue in shared memory;
//   B. when the queue holds 32 candidates they form a dense chunk that is tested against the REST of the list
//      (entries 32.., the warps sharing the groups as in nms_wide_kernel) and resolved by warp 0.
// The order of the decisions is the candidate order, as in the reference's loop: boxes are only kept in stage B (or in
// the plain chunks before 32 are kept), queue entries are drained first-in first-out, and a queued candidate has met
// every box kept before it -- the prefix in stage A, the rest in stage B.
constexpr int kStW = 4;
constexpr int kStF = 32;
constexpr int kStQ = 32 + kStW * 32;
__host__ __device__ constexpr int nms_staged_smem(int M) {
    return ((M + kWidePad + 32) * 20 + M * 4 + kStQ * 24 + 15) & ~15;
}

__device__ __forceinline__ bool box_in_unit(float4 b) {
    return b.x >= 0.f && b.y >= 0.f && b.z <= 1.f && b.w <= 1.f && b.x <= 1.f && b.y <= 1.f && b.z >= 0.f && b.w >= 0.f;
}

__global__ void __launch_bounds__(kStW * 32)
nms_staged_kernel(const __grid_constant__ NmsParams p) {
    constexpr int W = kStW;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned s_dead[W], s_unc[W];
    __shared__ unsigned s_step[2][W];
    __shared__ int s_cnt[W];
    __shared__ int s_count, s_kunit;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int seg = blockIdx.x;
    if (p.only_flagged && !p.only_flagged[seg]) return;
    float4* s_kbox = reinterpret_cast<float4*>(smem);
    float4* s_cbox = s_kbox + p.M + kWidePad;
    float4* q_box = s_cbox + 32;
    float* s_cvol = reinterpret_cast<float*>(q_box + kStQ);
    float* s_kvol = s_cvol + 32;
    float* s_kt = s_kvol + p.M;
    float* q_score = s_kt + p.M + kWidePad;
    int* q_pos = reinterpret_cast<int*>(q_score + kStQ);
    for (int i = threadIdx.x; i < p.M + kWidePad; i += W * 32) {
        s_kbox[i] = make_float4(2.f, 2.f, -1.f, -1.f);
        s_kt[i] = __int_as_float(0x7f800000);
    }
    if (threadIdx.x == 0) { s_count = 0; s_kunit = 1; }
    __syncthreads();
    const bool fastmode = (p.mode == RONK_NMS_MIN) && (p.thr > 0.f);
    const bool zero_supp = !(0.f < p.thr);
    const size_t in0 = (size_t)seg * p.K, out0 = (size_t)seg * p.M;
    int count = 0, kunit = 1;      // kept boxes so far; all of them inside [0, 1]^2
    int q_n = 0, c_next = 0;       // queue length; next candidate to load

    // One chunk (the same 32 candidates in every warp, lane = candidate) against kept[from .. count), resolved and
    // appended by warp 0.  Ends with a barrier; count / kunit are refreshed from shared memory.
    auto process_chunk = [&](float4 box, float score, int pos, bool valid, int from) {
        const float vol = (box.w - box.y) * (box.z - box.x);
        bool dead = !valid;
        bool unclear = false;
        const bool unit = fastmode && kunit && __all_sync(full, box_in_unit(box));
        if (unit) {
            const float tj = p.thr * vol;
            const float tolj = tj * 1e-6f;
            int par = 0;
            for (int s0 = from; s0 < count; s0 += W * 2 * kNmsPad) {
                const int i0 = s0 + warp * 2 * kNmsPad;
                if (i0 < count) {
#pragma unroll
                    for (int i = 0; i < 2 * kNmsPad; ++i) {
                        const float4 kb = s_kbox[i0 + i];
                        const float kt = s_kt[i0 + i];
                        const float h = __saturatef(fminf(box.z, kb.z) - fmaxf(box.x, kb.x));
                        const float w = __saturatef(fminf(box.w, kb.w) - fmaxf(box.y, kb.y));
                        const float d = h * w - fminf(tj, kt);
                        dead |= d > 0.f;
                        unclear |= fabsf(d) <= tolj;
                    }
                }
                if (s0 + W * 2 * kNmsPad >= count) break;
                const unsigned dmi = __ballot_sync(full, dead);
                if (lane == 0) s_step[par][warp] = dmi;
                __syncthreads();
                unsigned dall = 0u;
#pragma unroll
                for (int w = 0; w < W; ++w) dall |= s_step[par][w];
                par ^= 1;
                if (dall == full) { dead = true; break; }
            }
            if (valid && !(vol > 0.f)) { dead = false; unclear = false; }
        } else {
            for (int i0 = from + warp * 8; i0 < count; i0 += W * 8) {
                const int lim = min(8, count - i0);
                for (int i = 0; i < lim; ++i)
                    dead = dead || suppresses(box, vol, s_kbox[i0 + i], s_kvol[i0 + i], p.mode, p.thr, zero_supp);
            }
        }
        const unsigned dm = __ballot_sync(full, dead), um = __ballot_sync(full, unclear && valid);
        if (lane == 0) { s_dead[warp] = dm; s_unc[warp] = um; }
        __syncthreads();
        if (warp == 0) {
            unsigned dall = 0u, uall = 0u;
#pragma unroll
            for (int w = 0; w < W; ++w) { dall |= s_dead[w]; uall |= s_unc[w]; }
            dead = (dall >> lane) & 1u;
            if (unit && valid && !(vol > 0.f)) dead = false;
            if (uall) {
                dead = !valid;
                for (int i = from; i < count; ++i)
                    dead |= suppresses(box, vol, s_kbox[i], s_kvol[i], p.mode, p.thr, zero_supp);
            }
            s_cbox[lane] = box;
            s_cvol[lane] = vol;
            __syncwarp();
            unsigned alive_mask = __ballot_sync(full, !dead);
            unsigned kept = 0u;
            int room = p.M - count;
            while (alive_mask && room > 0) {
                const int i = __ffs(alive_mask) - 1;
                kept |= 1u << i;
                --room;
                const float4 kb = s_cbox[i];
                const float kv = s_cvol[i];
                if (lane > i && !dead) dead = suppresses(box, vol, kb, kv, p.mode, p.thr, zero_supp);
                alive_mask = __ballot_sync(full, !dead) & ~((2u << i) - 1u);
            }
            const bool mine = (kept >> lane) & 1u;
            if (mine) {
                const int r = count + __popc(kept & ((1u << lane) - 1u));
                s_kbox[r] = box;
                s_kvol[r] = vol;
                s_kt[r] = (vol > 0.f) ? p.thr * vol : __int_as_float(0x7f800000);
                p.out_scores[out0 + r] = score;
                p.out_boxes[out0 + r] = box;
                if (p.out_idx) p.out_idx[out0 + r] = pos;
            }
            const bool ok = __all_sync(full, !mine || box_in_unit(box));
            if (lane == 0) {
                s_count = count + __popc(kept);
                if (!ok) s_kunit = 0;
            }
        }
        __syncthreads();
        count = s_count;
        kunit = s_kunit;
    };

    while (count < p.M) {
        if (count < kStF) {
            // ---- plain chunks until the fixed prefix is complete
            if (c_next >= p.K) break;
            const int j = c_next + lane;
            const bool valid = j < p.K;
            int pos = -1;
            float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
            float score = 0.f;
            if (valid) {
                pos = p.order ? p.order[in0 + j] : j;
                box = p.boxes[in0 + pos];
                score = p.scores[in0 + pos];
            }
            c_next += 32;
            process_chunk(box, score, pos, valid, 0);
            continue;
        }
        if (q_n < 32 && c_next < p.K) {
            // ---- stage A: warp w filters candidates [c_next + 32 w, +32) through the prefix kept[0 .. 32)
            const int j = c_next + 32 * warp + lane;
            const bool valid = j < p.K;
            int pos = -1;
            float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
            float score = 0.f;
            if (valid) {
                pos = p.order ? p.order[in0 + j] : j;
                box = p.boxes[in0 + pos];
                score = p.scores[in0 + pos];
            }
            const float vol = (box.w - box.y) * (box.z - box.x);
            bool dead = !valid;
            if (fastmode && kunit && __all_sync(full, box_in_unit(box))) {
                const float tj = p.thr * vol;
                const float tolj = tj * 1e-6f;
                bool unclear = false;
#pragma unroll 8
                for (int i = 0; i < kStF; ++i) {
                    const float4 kb = s_kbox[i];
                    const float kt = s_kt[i];
                    const float h = __saturatef(fminf(box.z, kb.z) - fmaxf(box.x, kb.x));
                    const float w = __saturatef(fminf(box.w, kb.w) - fmaxf(box.y, kb.y));
                    const float d = h * w - fminf(tj, kt);
                    dead |= d > 0.f;
                    unclear |= fabsf(d) <= tolj;
                }
                if (valid && !(vol > 0.f)) { dead = false; unclear = false; }     // an empty candidate overlaps nothing
                if (unclear && valid) {                                           // inside the band: the exact quotient decides
                    dead = false;
                    for (int i = 0; i < kStF; ++i)
                        dead = dead || suppresses(box, vol, s_kbox[i], s_kvol[i], p.mode, p.thr, zero_supp);
                }
            } else if (valid) {
                for (int i = 0; i < kStF; ++i)
                    dead = dead || suppresses(box, vol, s_kbox[i], s_kvol[i], p.mode, p.thr, zero_supp);
            }
            const unsigned am = __ballot_sync(full, !dead);
            if (lane == 0) s_cnt[warp] = __popc(am);
            __syncthreads();
            int off = q_n, tot = 0;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                const int c = s_cnt[w];
                off += (w < warp) ? c : 0;
                tot += c;
            }
            if (!dead) {
                const int r = off + __popc(am & ((1u << lane) - 1u));
                q_box[r] = box;
                q_score[r] = score;
                q_pos[r] = pos;
            }
            __syncthreads();
            q_n += tot;
            c_next += 32 * W;
            continue;
        }
        if (q_n == 0) break;
        // ---- stage B: the oldest (up to) 32 queued candidates as one chunk against kept[32 .. count)
        const int nq = min(32, q_n);
        {
            const bool valid = lane < nq;
            float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
            float score = 0.f;
            int pos = -1;
            if (valid) { box = q_box[lane]; score = q_score[lane]; pos = q_pos[lane]; }
            process_chunk(box, score, pos, valid, kStF);
        }
        // the rest of the queue moves to the front (at most W * 32 entries: one per thread)
        const int rem = q_n - nq;
        {
            const int i = threadIdx.x;
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            float sc = 0.f;
            int ps = -1;
            if (i < rem) { b = q_box[nq + i]; sc = q_score[nq + i]; ps = q_pos[nq + i]; }
            __syncthreads();
            if (i < rem) { q_box[i] = b; q_score[i] = sc; q_pos[i] = ps; }
            __syncthreads();
        }
        q_n = rem;
    }
    if (p.out_short && threadIdx.x == 0) p.out_short[seg] = count < p.M ? 1 : 0;
    for (int r = count + (int)threadIdx.x; r < p.M; r += W * 32) {
        p.out_scores[out0 + r] = 0.f;
        p.out_boxes[out0 + r] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.out_idx) p.out_idx[out0 + r] = -1;
    }
}

// stable descending order of every row (tf.nn.top_k(k = row length), bboxes.py:179-180)
struct RowSrc {
    const float* g;
    __device__ __forceinline__ u64 get(int i) const {
        unsigned u = __float_as_uint(g[i] + 0.f);          // -0 -> +0: the two zeros tie (lower index first)
        u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
        return ((u64)u << 32) | (u64)(0xffffffffu - (unsigned)i);
    }
};

__global__ void __launch_bounds__(kTopkThreads)
row_order_kernel(const float* __restrict__ scores, int K, int* __restrict__ order) {
    extern __shared__ __align__(128) unsigned char smem[];
    u64* s_sort = reinterpret_cast<u64*>(smem);
    __shared__ unsigned s_hist[256];
    __shared__ int s_ctl[4];
    const size_t row = blockIdx.x;
    RowSrc src{scores + row * K};
    block_topk_sorted(src, K, K, s_hist, s_ctl, s_sort);
    for (int r = threadIdx.x; r < K; r += kTopkThreads)
        order[row * K + r] = (int)(0xffffffffu - (unsigned)(s_sort[r] & 0xffffffffull));
}

}  // namespace ronk

using namespace ronk;

extern "C" size_t ronk_nms_workspace_bytes(int S, int K) {
    if (S < 1 || K < 1) return 0;
    return (size_t)S * K * sizeof(int);
}

static int nms_launch(const float* scores, const float* boxes, int S, int K, float nms_threshold, int keep_top_k, int mode,
                      int assume_sorted, const int32_t* only_flagged, int32_t* out_short, float* out_scores,
                      float* out_boxes, int32_t* out_idx, void* ws, void* stream);

extern "C" int ronk_nms_batch(const float* scores, const float* boxes, int S, int K, float nms_threshold,
                              int keep_top_k, int mode, int assume_sorted, float* out_scores, float* out_boxes,
                              int32_t* out_idx, void* ws, void* stream) {
    return nms_launch(scores, boxes, S, K, nms_threshold, keep_top_k, mode, assume_sorted, nullptr, nullptr, out_scores,
                      out_boxes, out_idx, ws, stream);
}

// Sorted rows only.  only_flagged (or NULL): rows whose flag is 0 are skipped, their outputs stay as they are.
// out_short (or NULL): 1 for a row whose K candidates were all walked with fewer than keep_top_k kept -- with more
// candidates than K available the greedy loop of the reference would have gone on (two-tier top-k, see ronk.h).
extern "C" int ronk_nms_batch_tiered(const float* scores, const float* boxes, int S, int K, float nms_threshold,
                                     int keep_top_k, int mode, const int32_t* only_flagged, int32_t* out_short,
                                     float* out_scores, float* out_boxes, int32_t* out_idx, void* stream) {
    return nms_launch(scores, boxes, S, K, nms_threshold, keep_top_k, mode, 1, only_flagged, out_short, out_scores, out_boxes,
                      out_idx, nullptr, stream);
}

static int nms_launch(const float* scores, const float* boxes, int S, int K, float nms_threshold, int keep_top_k, int mode,
                      int assume_sorted, const int32_t* only_flagged, int32_t* out_short, float* out_scores,
                      float* out_boxes, int32_t* out_idx, void* ws, void* stream) {
    RONK_REQUIRE(scores && boxes && out_scores && out_boxes, RONK_EINVAL, "ronk_nms_batch: NULL argument");
    RONK_REQUIRE(S >= 1 && K >= 1 && keep_top_k >= 1, RONK_EINVAL, "ronk_nms_batch: S, K, keep_top_k must be >= 1");
    RONK_REQUIRE(mode == RONK_NMS_MIN || mode == RONK_NMS_UNION, RONK_EINVAL, "unknown mode to use for nms.");
    RONK_REQUIRE(((uintptr_t)boxes % 16) == 0 && ((uintptr_t)out_boxes % 16) == 0, RONK_EINVAL,
                 "ronk_nms_batch: box pointers must be 16-byte aligned");
    RONK_REQUIRE(assume_sorted || ws, RONK_EINVAL, "ronk_nms_batch: workspace required unless assume_sorted");
    RONK_REQUIRE(assume_sorted || K <= 16384, RONK_ELIMIT, "ronk_nms_batch: unsorted rows support K <= 16384");
    RONK_REQUIRE(keep_top_k <= 2048, RONK_ELIMIT, "ronk_nms_batch: keep_top_k <= 2048");
    cudaStream_t st = (cudaStream_t)stream;
    NmsParams p;
    p.scores = scores;
    p.boxes = (const float4*)boxes;
    p.order = nullptr;
    p.S = S; p.K = K; p.M = keep_top_k; p.mode = mode; p.thr = nms_threshold;
    p.out_scores = out_scores;
    p.out_boxes = (float4*)out_boxes;
    p.out_idx = out_idx;
    p.only_flagged = only_flagged;
    p.out_short = out_short;
    if (!assume_sorted) {
        int P = 1;
        while (P < K) P <<= 1;
        size_t smem = (size_t)P * sizeof(u64);
        if (smem > 48 * 1024)
            RONK_CUDA(cudaFuncSetAttribute(row_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        row_order_kernel<<<S, kTopkThreads, smem, st>>>(scores, K, (int*)ws);
        RONK_LAUNCHED();
        p.order = (const int*)ws;
    }
    size_t smem = (size_t)kNmsWarps * (size_t)nms_smem_per_warp(keep_top_k);
    if (smem > 48 * 1024)
        RONK_CUDA(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // few segments with long lists: staged CTA-per-segment kernel (nms_staged_kernel).  RONK_NMS_STAGED=0/1 overrides.
    int staged = (K >= 256 && S <= 16 * 148 && keep_top_k > kStF) ? 1 : 0;
    if (const char* e = getenv("RONK_NMS_STAGED")) staged = atoi(e) && keep_top_k > kStF;
    if (staged) {
        const size_t smem_s = (size_t)nms_staged_smem(keep_top_k);
        if (smem_s > 48 * 1024)
            RONK_CUDA(cudaFuncSetAttribute(nms_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
        nms_staged_kernel<<<S, kStW * 32, smem_s, st>>>(p);
        RONK_LAUNCHED();
        return RONK_OK;
    }
    // few segments with long lists: a CTA per segment (nms_wide_kernel).  RONK_NMS_WIDE=0 / 4 / 8 overrides.
    int wide = (K >= 1024 && S <= 16 * 148) ? 4 : 0;
    if (const char* e = getenv("RONK_NMS_WIDE")) wide = atoi(e);
    if (wide == 4 || wide == 8) {
        const size_t smem_w = (size_t)nms_wide_smem(keep_top_k);
        if (wide == 4) {
            if (smem_w > 48 * 1024)
                RONK_CUDA(cudaFuncSetAttribute(nms_wide_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));
            nms_wide_kernel<4><<<S, 128, smem_w, st>>>(p);
        } else {
            if (smem_w > 48 * 1024)
                RONK_CUDA(cudaFuncSetAttribute(nms_wide_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));
            nms_wide_kernel<8><<<S, 256, smem_w, st>>>(p);
        }
        RONK_LAUNCHED();
        return RONK_OK;
    }
    nms_kernel<<<(S + kNmsWarps - 1) / kNmsWarps, kNmsWarps * 32, smem, st>>>(p);
    RONK_LAUNCHED();
    return RONK_OK;
}
