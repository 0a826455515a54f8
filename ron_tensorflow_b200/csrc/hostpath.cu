// Host-buffer path of match + encode.  The targets of a batch are 28 B per anchor (38 MB at batch 64), and 16 B of
// them are the localisation row, which is non-zero for ~1 % of the anchors only.  Instead of copying the dense
// localisation tensor over PCIe every step, sparse_rows_kernel packs its non-zero rows (index + row) into a
// fixed-capacity packet; the host applies the packet to a pinned array that it keeps zero elsewhere (the rows of
// the previous packet are cleared first).  Labels and scores travel dense (a host-side scatter of the ~5 % non-zero
// labels was measured slower than the DMA engine writing all of them).  Same results in host memory, 2.2x fewer
// bytes over the bus.  If a packet overflows its capacity the caller copies the dense tensor for that step instead
// (ronk_host_rows_apply says so).
#include <string.h>

#include "common.cuh"

namespace ronk {

// packet: header int32[4] = {rows, 0, 0, 0} | idx int32[cap] | pad to 16 | rows float4[cap]
__host__ __device__ inline size_t packet_rows_offset(int cap) { return (16 + (size_t)cap * 4 + 15) & ~(size_t)15; }

__global__ void __launch_bounds__(256)
sparse_rows_kernel(const float4* __restrict__ rows, long long T, int cap, unsigned char* __restrict__ packet) {
    int* header = reinterpret_cast<int*>(packet);
    int* idx = reinterpret_cast<int*>(packet + 16);
    float4* val = reinterpret_cast<float4*>(packet + packet_rows_offset(cap));
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < T; i0 += stride) {
        const long long i = i0 + lane;
        const float4 v = i < T ? rows[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        const bool nz = (v.x != 0.f) || (v.y != 0.f) || (v.z != 0.f) || (v.w != 0.f);     // NaN counts as non-zero
        const unsigned m = __ballot_sync(full, nz);
        int base = 0;
        if (lane == 0 && m) base = atomicAdd(header, __popc(m));
        base = __shfl_sync(full, base, 0);
        if (nz) {
            const int pos = base + __popc(m & ((1u << lane) - 1u));
            if (pos < cap) { idx[pos] = (int)i; val[pos] = v; }
        }
    }
}

}  // namespace ronk

using namespace ronk;

extern "C" size_t ronk_sparse_rows_packet_bytes(int cap) { return cap < 1 ? 0 : packet_rows_offset(cap) + (size_t)cap * 16; }

extern "C" int ronk_sparse_rows_pack(const float* rows, long long T, int cap, void* packet_dev, void* stream) {
    RONK_REQUIRE(rows && packet_dev && T >= 1 && T < (1ll << 31) && cap >= 1, RONK_EINVAL, "ronk_sparse_rows_pack: bad argument");
    RONK_REQUIRE(((uintptr_t)rows % 16) == 0 && ((uintptr_t)packet_dev % 16) == 0, RONK_EINVAL,
                 "ronk_sparse_rows_pack: rows and the packet must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    RONK_CUDA(cudaMemsetAsync(packet_dev, 0, 16, st));
    long long blocks = (T + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    sparse_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>((const float4*)rows, T, cap, (unsigned char*)packet_dev);
    RONK_LAUNCHED();
    return RONK_OK;
}

// Host side, no CUDA: zero the rows the previous packet of this array wrote, then write the new packet's rows.
// Returns 1 (and touches nothing) when the new packet overflowed its capacity: the caller must copy the dense
// tensor for this step and zero the array before the next sparse step.
extern "C" int ronk_host_rows_apply(const void* packet_host, const void* prev_packet_host, int cap, float* rows_host) {
    RONK_REQUIRE(packet_host && rows_host && cap >= 1, RONK_EINVAL, "ronk_host_rows_apply: bad argument");
    const unsigned char* pk = (const unsigned char*)packet_host;
    const int n = *(const int*)pk;
    if (n > cap) return 1;
    if (prev_packet_host) {
        const unsigned char* pp = (const unsigned char*)prev_packet_host;
        const int* idx = (const int*)(pp + 16);
        const int np = *(const int*)pp;
        for (int k = 0; k < np; ++k) memset(rows_host + 4 * (size_t)idx[k], 0, 16);
    }
    const int* idx = (const int*)(pk + 16);
    const float* val = (const float*)(pk + packet_rows_offset(cap));
    for (int k = 0; k < n; ++k) memcpy(rows_host + 4 * (size_t)idx[k], val + 4 * (size_t)k, 16);
    return RONK_OK;
}
