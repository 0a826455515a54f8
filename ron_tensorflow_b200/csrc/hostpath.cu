// Host-buffer path of match + encode.  The targets of a batch are 28 B per anchor (38 MB at batch 64), and 16 B of
// them are the localisation row, which is non-zero for ~1 % of the anchors only.  Instead of copying the dense
// localisation tensor over PCIe every step, sparse_rows_kernel packs its non-zero rows (index + row) into a
// fixed-capacity packet; the host applies the packet to a pinned array that it keeps zero elsewhere (the rows of
// the previous packet are cleared first).  Labels and scores travel dense (a host-side scatter of the ~5 % non-zero
// labels was measured slower than the DMA engine writing all of them).  Same results in host memory, 2.2x fewer
// bytes over the bus.  If a packet overflows its capacity the caller copies the dense tensor for that step instead
// (ronk_host_rows_apply says so).
#include <string.h>

#include <algorithm>

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

// Labels: 0 (background) for ~95 % of the anchors.  The non-zero ones travel as (index, value) pairs.
// packet: header int32[4] = {entries, a label outside int32, 0, 0} | idx int32[cap] | val int32[cap]
__global__ void __launch_bounds__(256)
sparse_labels_kernel(const long long* __restrict__ labels, long long T, int cap, unsigned char* __restrict__ packet) {
    int* header = reinterpret_cast<int*>(packet);
    int* idx = reinterpret_cast<int*>(packet + 16);
    int* val = idx + cap;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < T; i0 += stride) {
        const long long i = i0 + lane;
        const long long v = i < T ? labels[i] : 0ll;
        const bool nz = v != 0ll;
        if (nz && v != (long long)(int)v) header[1] = 1;
        const unsigned m = __ballot_sync(full, nz);
        int base = 0;
        if (lane == 0 && m) base = atomicAdd(header, __popc(m));
        base = __shfl_sync(full, base, 0);
        if (nz) {
            const int pos = base + __popc(m & ((1u << lane) - 1u));
            if (pos < cap) { idx[pos] = (int)i; val[pos] = (int)v; }
        }
    }
}

}  // namespace ronk

// ---- a small persistent pool of host threads for the expansion below (created on first use)
#include <condition_variable>
#include <mutex>
#include <thread>

namespace {

struct HostPool {
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::vector<std::thread> workers;
    void (*fn)(void*, int) = nullptr;
    void* arg = nullptr;
    int chunks = 0, next = 0, running = 0;
    unsigned long long generation = 0;
    bool stop = false;

    void worker() {
        unsigned long long seen = 0;
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv_work.wait(lk, [&] { return stop || generation != seen; });
            if (stop) return;
            seen = generation;
            while (next < chunks) {
                const int c = next++;
                lk.unlock();
                fn(arg, c);
                lk.lock();
            }
            if (--running == 0) cv_done.notify_all();
        }
    }
    void ensure(int n) {
        while ((int)workers.size() < n) workers.emplace_back([this] { worker(); });
    }
    // runs fn(arg, c) for c in [0, chunks) on `threads` threads (the caller is one of them)
    void run(void (*f)(void*, int), void* a, int nchunks, int threads) {
        if (threads <= 1 || nchunks <= 1) {
            for (int c = 0; c < nchunks; ++c) f(a, c);
            return;
        }
        std::unique_lock<std::mutex> lk(mu);
        ensure(threads - 1);
        fn = f; arg = a; chunks = nchunks; next = 0;
        running = (int)workers.size();
        ++generation;
        cv_work.notify_all();
        while (next < chunks) {
            const int c = next++;
            lk.unlock();
            f(a, c);
            lk.lock();
        }
        cv_done.wait(lk, [&] { return running == 0; });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv_work.notify_all();
        for (auto& t : workers) t.join();
    }
};

HostPool& pool() {
    static HostPool* p = new HostPool();      // leaked on purpose: no join at process exit
    return *p;
}

struct ApplyJob {
    int phase;                       // 0: zero what the previous packets wrote, 1: write the new packets
    const int* ridx; const float* rval; int rn;          // localisation rows of this phase
    float* rows;
    const int* lidx; const int* lval; int ln;            // labels of this phase
    long long* labels;
    int per_chunk;
};

void apply_chunk(void* a, int c) {
    const ApplyJob& j = *(const ApplyJob*)a;
    const int r0 = std::min(j.rn, c * j.per_chunk), r1 = std::min(j.rn, r0 + j.per_chunk);
    const int l0 = std::min(j.ln, c * j.per_chunk), l1 = std::min(j.ln, l0 + j.per_chunk);
    if (j.phase == 0) {
        for (int k = r0; k < r1; ++k) memset(j.rows + 4 * (size_t)j.ridx[k], 0, 16);
        for (int k = l0; k < l1; ++k) j.labels[j.lidx[k]] = 0;
    } else {
        for (int k = r0; k < r1; ++k) memcpy(j.rows + 4 * (size_t)j.ridx[k], j.rval + 4 * (size_t)k, 16);
        for (int k = l0; k < l1; ++k) j.labels[j.lidx[k]] = (long long)j.lval[k];
    }
}

}  // namespace

using namespace ronk;

extern "C" size_t ronk_sparse_labels_packet_bytes(int cap) { return cap < 1 ? 0 : 16 + (size_t)cap * 8; }

extern "C" int ronk_sparse_labels_pack(const int64_t* labels, long long T, int cap, void* packet_dev, void* stream) {
    RONK_REQUIRE(labels && packet_dev && T >= 1 && T < (1ll << 31) && cap >= 1, RONK_EINVAL, "ronk_sparse_labels_pack: bad argument");
    RONK_REQUIRE(((uintptr_t)packet_dev % 16) == 0, RONK_EINVAL, "ronk_sparse_labels_pack: the packet must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    RONK_CUDA(cudaMemsetAsync(packet_dev, 0, 16, st));
    long long blocks = (T + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    sparse_labels_kernel<<<(unsigned)blocks, 256, 0, st>>>((const long long*)labels, T, cap, (unsigned char*)packet_dev);
    RONK_LAUNCHED();
    return RONK_OK;
}

// Host side, no CUDA, `threads` host threads from a persistent pool: zero what the PREVIOUS packets of the same host
// arrays wrote (NULL: nothing), then write the new packets' localisation rows and labels.  The arrays must hold
// zeros elsewhere.  Returns a bit mask, and leaves the array in question untouched by the new packet: 1 = the
// localisation packet overflowed, 2 = the label packet overflowed or held a label outside int32 -- copy that tensor
// densely for this step and pass NULL as its previous packet next time (after zeroing / overwriting the array).
extern "C" int ronk_host_targets_apply(const void* loc_packet, const void* prev_loc_packet, int loc_cap, float* rows_host,
                                       const void* lab_packet, const void* prev_lab_packet, int lab_cap,
                                       int64_t* labels_host, int threads) {
    RONK_REQUIRE(loc_packet && rows_host && lab_packet && labels_host && loc_cap >= 1 && lab_cap >= 1, RONK_EINVAL,
                 "ronk_host_targets_apply: bad argument");
    const unsigned char* lp = (const unsigned char*)loc_packet;
    const unsigned char* bp = (const unsigned char*)lab_packet;
    const int rn = *(const int*)lp, ln = *(const int*)bp;
    const bool loc_bad = rn > loc_cap, lab_bad = ln > lab_cap || ((const int*)bp)[1] != 0;
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    ApplyJob j;
    j.rows = rows_host;
    j.labels = (long long*)labels_host;
    // phase 0: previous packets
    j.phase = 0;
    j.rn = 0; j.ln = 0;
    if (prev_loc_packet && !loc_bad) {
        const unsigned char* pp = (const unsigned char*)prev_loc_packet;
        j.rn = std::min(*(const int*)pp, loc_cap);
        j.ridx = (const int*)(pp + 16);
    }
    if (prev_lab_packet && !lab_bad) {
        const unsigned char* pp = (const unsigned char*)prev_lab_packet;
        j.ln = std::min(*(const int*)pp, lab_cap);
        j.lidx = (const int*)(pp + 16);
    }
    int most = std::max(j.rn, j.ln);
    if (most > 0) {
        j.per_chunk = std::max(1024, (most + threads * 4 - 1) / (threads * 4));
        pool().run(apply_chunk, &j, (most + j.per_chunk - 1) / j.per_chunk, threads);
    }
    // phase 1: new packets
    j.phase = 1;
    j.rn = loc_bad ? 0 : rn;
    j.ridx = (const int*)(lp + 16);
    j.rval = (const float*)(lp + packet_rows_offset(loc_cap));
    j.ln = lab_bad ? 0 : ln;
    j.lidx = (const int*)(bp + 16);
    j.lval = j.lidx + lab_cap;
    most = std::max(j.rn, j.ln);
    if (most > 0) {
        j.per_chunk = std::max(1024, (most + threads * 4 - 1) / (threads * 4));
        pool().run(apply_chunk, &j, (most + j.per_chunk - 1) / j.per_chunk, threads);
    }
    return (loc_bad ? 1 : 0) | (lab_bad ? 2 : 0);
}

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
