// Block-wide exact top-K of unique 64-bit keys: MSD radix select (8-bit digits) to find the
// K-th largest key, gather the K winners, bitonic sort descending in shared memory.
// Keys embed the element index in the low word ((0xffffffff - index), so that among equal
// scores the LOWER index is the LARGER key): the resulting order is exactly
// tf.nn.top_k(sorted=True) -- descending score, lower index first (SURVEY.md hard part 2/5).
#pragma once
#include "common.cuh"

namespace ronk {

constexpr int kTopkThreads = 256;

__device__ __forceinline__ int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// Src: struct with  __device__ u64 get(int i) const  for i in [0, n).
// s_hist: 256 unsigned; s_ctl: 4 ints; s_sort: pow2(K) u64.
// On return s_sort[0 .. P) is sorted descending with min(n, K) real keys then zeros.
template <class Src>
__device__ void block_topk_sorted(const Src& src, int n, int K, unsigned* s_hist, int* s_ctl, u64* s_sort) {
    const int tid = threadIdx.x;
    const int P = next_pow2(K);
    for (int i = tid; i < P; i += kTopkThreads) s_sort[i] = 0ull;
    if (tid == 0) s_ctl[3] = 0;
    u64 thr = 0ull;   // select keys >= thr
    if (n > K) {
        u64 prefix = 0ull, mask = 0ull;
        int need = K;
        for (int shift = 56; shift >= 0; shift -= 8) {
            for (int i = tid; i < 256; i += kTopkThreads) s_hist[i] = 0u;
            __syncthreads();
            for (int i0 = 0; i0 < n; i0 += kTopkThreads) {
                int i = i0 + tid;
                bool act = false;
                unsigned d = 0;
                if (i < n) {
                    u64 k = src.get(i);
                    act = (k & mask) == prefix;
                    d = (unsigned)(k >> shift) & 255u;
                }
                // warp-aggregated histogram: one atomic per distinct digit per warp
                unsigned amask = __ballot_sync(0xffffffffu, act);
                if (act) {
                    unsigned peers = __match_any_sync(amask, d);
                    if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&s_hist[d], (unsigned)__popc(peers));
                }
            }
            __syncthreads();
            if (tid < 32) {
                // digits 255 .. 0, lane l owns digits [255 - 8l - 7, 255 - 8l]; find the digit where the
                // count of keys with a larger digit is < need <= that count + hist[digit]
                unsigned c[8];
                unsigned sum = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) { c[q] = s_hist[255 - (tid * 8 + q)]; sum += c[q]; }
                unsigned incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (tid >= o) incl += v;
                }
                unsigned above = incl - sum;   // keys with digits larger than this lane's range
                if (above < (unsigned)need && (unsigned)need <= incl) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if (above < (unsigned)need && (unsigned)need <= above + c[q]) {
                            s_ctl[0] = 255 - (tid * 8 + q);
                            s_ctl[1] = need - (int)above;
                            s_ctl[2] = (c[q] == (unsigned)need - above) ? 1 : 0;
                            above = 0xffffffffu;   // stop
                        } else if (above != 0xffffffffu) {
                            above += c[q];
                        }
                    }
                }
            }
            __syncthreads();
            int d = s_ctl[0];
            need = s_ctl[1];
            int whole = s_ctl[2];
            prefix |= (u64)d << shift;
            mask |= 255ull << shift;
            if (whole) break;
        }
        thr = prefix;
    }
    __syncthreads();
    for (int i = tid; i < n; i += kTopkThreads) {
        u64 k = src.get(i);
        if (k >= thr && k != 0ull) {
            int pos = atomicAdd(&s_ctl[3], 1);
            if (pos < P) s_sort[pos] = k;
        }
    }
    __syncthreads();
    // bitonic sort, descending
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

}  // namespace ronk
