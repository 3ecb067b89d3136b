// Error reporting / bookkeeping entry points of the avr_b200 C-ABI.
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include "common.cuh"

namespace avr {

static thread_local char g_error[512] = "";
static std::atomic<int64_t> g_launches{0};   // process-wide statistic only

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char* where) {
    snprintf(g_error, sizeof(g_error), "%s: CUDA error %d (%s)", where, (int)e, cudaGetErrorString(e));
    return (int)e;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace avr

extern "C" int avr_abi_version(void) { return AVR_B200_ABI_VERSION; }

extern "C" const char* avr_last_error(void) { return avr::g_error; }

extern "C" int64_t avr_launch_count(int reset) {
    return reset ? avr::g_launches.exchange(0) : avr::g_launches.load();
}
