// Library-wide pieces of the C ABI: version, thread-local error text, launch counter.
#include "common.cuh"

namespace ronk {

static thread_local std::string t_error;
std::atomic<long long> g_launches{0};

void set_error(const std::string& msg) { t_error = msg; }

int cuda_fail(cudaError_t e, const char* what) {
    t_error = std::string(what) + ": " + cudaGetErrorName(e) + ": " + cudaGetErrorString(e);
    return RONK_ECUDA;
}

}  // namespace ronk

extern "C" int ronk_version(void) { return RONK_VERSION; }
extern "C" const char* ronk_last_error(void) { return ronk::t_error.c_str(); }
extern "C" long long ronk_launch_count(void) { return ronk::g_launches.load(std::memory_order_relaxed); }

// 0 when `stream` is not being captured into a CUDA graph, else the id of the capture (shared by every stream forked
// into it): the Python glue keys capture-time workspaces by it (a buffer allocated during a capture belongs to that graph).
extern "C" int ronk_stream_capture_id(void* stream, unsigned long long* out_id) {
    RONK_REQUIRE(out_id != nullptr, RONK_EINVAL, "ronk_stream_capture_id: out_id is NULL");
    cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
    unsigned long long id = 0;
    RONK_CUDA(cudaStreamGetCaptureInfo((cudaStream_t)stream, &status, &id));
    *out_id = status == cudaStreamCaptureStatusActive ? id : 0ull;
    return RONK_OK;
}
