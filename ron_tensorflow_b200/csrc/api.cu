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
