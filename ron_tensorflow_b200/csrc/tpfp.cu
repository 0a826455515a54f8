// TP/FP matching of detections against ground truth, one warp per (image, class).
//
// Reference: tf_extended/bboxes.py:316-404 (bboxes_matching), :407-450 (bboxes_matching_batch),
// :527-554 (bboxes_jaccard).  Spec: SURVEY.md Appendix A.8.  Bit-exact against the oracle.
//
// The reference walks the detections one by one because a ground-truth box can be matched only
// once.  Only that "already matched" bit is sequential: the jaccard arg-max of a detection does
// not depend on the others.  So every lane takes one detection of a 32-wide chunk, scans the
// image's ground truth (staged once per CTA in shared memory) for its first arg-max, and the
// chunk is then resolved in one step: a detection sees its GT as "existing" if an earlier chunk
// matched it (bitmask in shared memory) or an earlier lane of this chunk does (match_any +
// lane mask).  M detections cost ceil(M/32) steps instead of M.
#include "common.cuh"

namespace ronk {

constexpr int kTpfpWarps = 8;

struct TpfpParams {
    const float* det_scores;
    const float4* det_boxes;
    int B, C, M, Gmax;
    const long long* glabels;
    const float4* gboxes;
    const long long* gdiff;
    float thr;
    long long* out_n_gt;
    uint8_t* out_tp;
    uint8_t* out_fp;
};

__global__ void __launch_bounds__(kTpfpWarps * 32)
tpfp_kernel(const __grid_constant__ TpfpParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    float4* s_gbox = reinterpret_cast<float4*>(smem);                 // [Gmax]
    float* s_garea = reinterpret_cast<float*>(s_gbox + p.Gmax);       // [Gmax]
    int* s_glab = reinterpret_cast<int*>(s_garea + p.Gmax);           // [Gmax]
    int* s_gdiff = s_glab + p.Gmax;                                   // [Gmax]
    unsigned* s_match = reinterpret_cast<unsigned*>(s_gdiff + p.Gmax);   // [warps][ceil(Gmax/32)]
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int CM = p.C - 1;
    const int words = (p.Gmax + 31) / 32;
    for (int g = threadIdx.x; g < p.Gmax; g += blockDim.x) {
        float4 gb = p.gboxes[(size_t)b * p.Gmax + g];
        s_gbox[g] = gb;
        s_garea[g] = (gb.z - gb.x) * (gb.w - gb.y);
        long long l = p.glabels[(size_t)b * p.Gmax + g];
        s_glab[g] = (l < -2147483647ll || l > 2147483647ll) ? -2147483647 : (int)l;
        s_gdiff[g] = p.gdiff[(size_t)b * p.Gmax + g] != 0;
    }
    for (int i = threadIdx.x; i < kTpfpWarps * words; i += blockDim.x) s_match[i] = 0u;
    __syncthreads();
    const int ci = blockIdx.x * kTpfpWarps + warp;   // class index 0..C-2
    if (ci >= CM) return;
    const int label = ci + 1;
    unsigned* match = s_match + warp * words;

    // n_gbboxes = #(glabel == c and not difficult)   (bboxes.py:344-345)
    int cnt = 0;
    for (int g = lane; g < p.Gmax; g += 32) cnt += (s_glab[g] == label && !s_gdiff[g]) ? 1 : 0;
    cnt = __reduce_add_sync(full, cnt);
    const size_t seg = (size_t)b * CM + ci;
    if (lane == 0) p.out_n_gt[seg] = cnt;

    for (int i0 = 0; i0 < p.M; i0 += 32) {
        const int i = i0 + lane;
        const bool valid = i < p.M;
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) r = p.det_boxes[seg * p.M + i];
        const float rarea = (r.z - r.x) * (r.w - r.y);
        // jaccard vs every GT, masked by class; first arg-max.  GT boxes of another class have
        // jaccard * 0 = 0 and every jaccard is >= 0, so starting from (0, GT 0) and visiting only
        // this class's boxes in ascending order with a strict '>' gives the same first arg-max.
        float best = 0.f;
        int idx = 0;
        for (int g0 = 0; g0 < p.Gmax; g0 += 32) {
            unsigned mine = __ballot_sync(full, g0 + lane < p.Gmax && s_glab[min(g0 + lane, p.Gmax - 1)] == label);
            while (mine) {
                const int g = g0 + __ffs(mine) - 1;
                mine &= mine - 1;
                float4 gb = s_gbox[g];
                float h = fmaxf(fminf(gb.z, r.z) - fmaxf(gb.x, r.x), 0.f);
                float w = fmaxf(fminf(gb.w, r.w) - fmaxf(gb.y, r.y), 0.f);
                float inter = h * w;
                float uni = (-inter + s_garea[g]) + rarea;
                float jac = div_overlap(inter, uni);   // safe_divide: inter > 0 implies uni > 0
                if (jac > best) { best = jac; idx = g; }
            }
        }
        const bool is_match = best > p.thr;
        const bool nd = !s_gdiff[idx];
        const bool marks = valid && nd && is_match;            // this detection marks its GT (bboxes.py:377-380)
        const unsigned markers = __ballot_sync(full, marks);
        const unsigned same = __match_any_sync(full, idx);
        const bool earlier = (same & markers & ((1u << lane) - 1u)) != 0u;
        const bool existing = (((match[idx >> 5] >> (idx & 31)) & 1u) != 0u) || earlier;
        const bool tp = nd && is_match && !existing;
        const bool fp = nd && (existing || !is_match);
        if (valid) {
            p.out_tp[seg * p.M + i] = tp ? 1 : 0;
            p.out_fp[seg * p.M + i] = fp ? 1 : 0;
        }
        __syncwarp();
        if (marks) atomicOr(&match[idx >> 5], 1u << (idx & 31));
        __syncwarp();
    }
}

// ---- device-resident TP/FP records (tf_extended/metrics.py:133-206 without the per-batch trip to the host)
// A record is one detection that survives the reference's filter (:167-175): score > min_score and (tp or fp), packed
// as (score bits << 32) | (class index << 8) | fp << 1 | tp.  One batch is appended in (class, image, rank) order --
// per class that is the order of the reference's tf.boolean_mask over the flattened [B, M] tensors -- behind the
// records of the earlier batches.  totals = {count before, count after, overflow flag, -}: the host flips the roles of
// the first two slots from call to call, so no call reads a value another block of the same call writes.
constexpr int kRecTile = 2048;

struct RecParams {
    const float* scores;        // [B, CM, M]
    const uint8_t* tp;
    const uint8_t* fp;
    const long long* n_gt;      // [B, CM]
    int B, CM, M;
    float min_score;
    long long n;                // CM * B * M
    int* tile_counts;
    u64* rec;
    int cap;
    int* totals;
    int slot_in, slot_out;
    long long* n_gt_acc;        // [CM]
    int* seg_counts;            // [CM] records of this call per class (zeroed before the launch), or NULL
};

__device__ __forceinline__ bool rec_keep(const RecParams& p, long long e, size_t* src) {
    const long long per_class = (long long)p.B * p.M;
    const int c = (int)(e / per_class);
    const long long r = e - (long long)c * per_class;
    const int b = (int)(r / p.M), m = (int)(r - (long long)b * p.M);
    *src = ((size_t)b * p.CM + c) * p.M + m;
    return p.scores[*src] > p.min_score && (p.tp[*src] | p.fp[*src]);
}

__global__ void __launch_bounds__(256)
records_count_kernel(const __grid_constant__ RecParams p) {
    __shared__ int s_w[8];
    const long long base = (long long)blockIdx.x * kRecTile;
    int c = 0;
    const long long per_class = (long long)p.B * p.M;
    for (int k = threadIdx.x; k < kRecTile; k += 256) {
        const long long e = base + k;
        size_t src;
        const bool keep = e < p.n && rec_keep(p, e, &src);
        c += keep ? 1 : 0;
        if (p.seg_counts) {
            // per-class totals of this call: the 32 elements of a warp step nearly always share one class
            const int cls = keep ? (int)(e / per_class) : -1;
            const unsigned peers = __match_any_sync(0xffffffffu, cls);
            if (keep && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(p.seg_counts + cls, __popc(peers));
        }
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += s_w[w];
        p.tile_counts[blockIdx.x] = t;
    }
    if (blockIdx.x == 0)                                   // ground-truth counts of the batch (metrics.py:177,183)
        for (int cc = threadIdx.x; cc < p.CM; cc += 256) {
            long long t = 0;
            for (int b = 0; b < p.B; ++b) t += p.n_gt[(size_t)b * p.CM + cc];
            p.n_gt_acc[cc] += t;
        }
}

__global__ void __launch_bounds__(256)
records_write_kernel(const __grid_constant__ RecParams p) {
    __shared__ int s_w[8];
    __shared__ int s_base;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int c = 0;
    for (int t = threadIdx.x; t < (int)blockIdx.x; t += 256) c += p.tile_counts[t];
    c = __reduce_add_sync(full, c);
    if (lane == 0) s_w[warp] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += s_w[w];
        const int before = p.totals[p.slot_in];
        s_base = before + t;
        if (blockIdx.x == gridDim.x - 1) {
            const int after = before + t + p.tile_counts[blockIdx.x];
            p.totals[p.slot_out] = after < p.cap ? after : p.cap;
            if (after > p.cap) p.totals[2] = 1;
        }
    }
    __syncthreads();
    int pos = s_base;
    const long long base = (long long)blockIdx.x * kRecTile;
    const long long per_class = (long long)p.B * p.M;
    for (int k0 = 0; k0 < kRecTile; k0 += 256) {
        const long long e = base + k0 + threadIdx.x;
        size_t src = 0;
        const bool keep = e < p.n && rec_keep(p, e, &src);
        const unsigned m = __ballot_sync(full, keep);
        __syncthreads();
        if (lane == 0) s_w[warp] = __popc(m);
        __syncthreads();
        int before = 0, all = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            before += (w < warp) ? s_w[w] : 0;
            all += s_w[w];
        }
        if (keep) {
            const int o = pos + before + __popc(m & ((1u << lane) - 1u));
            if (o < p.cap) {
                const unsigned cls = (unsigned)(e / per_class);
                p.rec[o] = ((u64)__float_as_uint(p.scores[src]) << 32) | (u64)((cls << 8) | (p.fp[src] ? 2u : 0u) | (p.tp[src] ? 1u : 0u));
            }
        }
        pos += all;
    }
}

}  // namespace ronk

using namespace ronk;

extern "C" size_t ronk_tpfp_records_workspace_bytes(int B, int C, int M) {
    if (B < 1 || C < 2 || M < 1) return 0;
    const long long n = (long long)(C - 1) * B * M;
    return (size_t)((n + kRecTile - 1) / kRecTile) * 4;
}

extern "C" int ronk_tpfp_records_append(const float* det_scores, const uint8_t* tp, const uint8_t* fp, const int64_t* n_gt,
                                        int B, int C, int M, float min_score, uint64_t* records, int capacity,
                                        int32_t* totals, int call_parity, int64_t* n_gt_acc, int32_t* seg_counts, void* ws,
                                        void* stream) {
    RONK_REQUIRE(det_scores && tp && fp && n_gt && records && totals && n_gt_acc && ws, RONK_EINVAL,
                 "ronk_tpfp_records_append: NULL argument");
    RONK_REQUIRE(B >= 1 && C >= 2 && C <= (1 << 20) && M >= 1 && capacity >= 1, RONK_EINVAL, "ronk_tpfp_records_append: bad sizes");
    RecParams p;
    p.scores = det_scores; p.tp = tp; p.fp = fp; p.n_gt = (const long long*)n_gt;
    p.B = B; p.CM = C - 1; p.M = M;
    p.min_score = min_score;
    p.n = (long long)p.CM * B * M;
    RONK_REQUIRE(p.n < (1ll << 31), RONK_ELIMIT, "ronk_tpfp_records_append: batch too large");
    p.tile_counts = (int*)ws;
    p.rec = (u64*)records;
    p.cap = capacity;
    p.totals = totals;
    p.slot_in = call_parity & 1;
    p.slot_out = p.slot_in ^ 1;
    p.n_gt_acc = (long long*)n_gt_acc;
    p.seg_counts = seg_counts;
    if (seg_counts) RONK_CUDA(cudaMemsetAsync(seg_counts, 0, (size_t)p.CM * 4, (cudaStream_t)stream));
    const unsigned tiles = (unsigned)((p.n + kRecTile - 1) / kRecTile);
    cudaStream_t st = (cudaStream_t)stream;
    records_count_kernel<<<tiles, 256, 0, st>>>(p);
    RONK_LAUNCHED();
    records_write_kernel<<<tiles, 256, 0, st>>>(p);
    RONK_LAUNCHED();
    return RONK_OK;
}

extern "C" int ronk_tpfp_match(const float* det_scores, const float* det_boxes, int B, int C, int M,
                               const int64_t* glabels, const float* gboxes, const int64_t* gdifficults, int Gmax,
                               float matching_threshold, int64_t* out_n_gt, uint8_t* out_tp, uint8_t* out_fp,
                               void* stream) {
    RONK_REQUIRE(det_boxes && glabels && gboxes && gdifficults && out_n_gt && out_tp && out_fp, RONK_EINVAL,
                 "ronk_tpfp_match: NULL argument");
    RONK_REQUIRE(B >= 1 && C >= 2 && M >= 1 && Gmax >= 1, RONK_EINVAL, "ronk_tpfp_match: bad sizes");
    RONK_REQUIRE(B <= 65535, RONK_ELIMIT, "ronk_tpfp_match: B <= 65535 per call");
    RONK_REQUIRE(((uintptr_t)det_boxes % 16) == 0 && ((uintptr_t)gboxes % 16) == 0, RONK_EINVAL,
                 "ronk_tpfp_match: box pointers must be 16-byte aligned");
    size_t smem = (size_t)Gmax * (16 + 4 + 4 + 4) + (size_t)kTpfpWarps * ((Gmax + 31) / 32) * 4;
    RONK_REQUIRE(smem <= 200 * 1024, RONK_ELIMIT, "ronk_tpfp_match: Gmax too large");
    if (smem > 48 * 1024)
        RONK_CUDA(cudaFuncSetAttribute(tpfp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TpfpParams p;
    p.det_scores = det_scores;
    p.det_boxes = (const float4*)det_boxes;
    p.B = B; p.C = C; p.M = M; p.Gmax = Gmax;
    p.glabels = (const long long*)glabels;
    p.gboxes = (const float4*)gboxes;
    p.gdiff = (const long long*)gdifficults;
    p.thr = matching_threshold;
    p.out_n_gt = (long long*)out_n_gt;
    p.out_tp = out_tp;
    p.out_fp = out_fp;
    dim3 grid((C - 1 + kTpfpWarps - 1) / kTpfpWarps, B);
    tpfp_kernel<<<grid, kTpfpWarps * 32, smem, (cudaStream_t)stream>>>(p);
    RONK_LAUNCHED();
    return RONK_OK;
}
