// Error reporting and device queries shared by all entry points of libzsv_b200.so.
#include <atomic>
#include <mutex>
#include <string.h>

#include <stdlib.h>

#include "zsv_internal.h"

namespace zsv {

static thread_local char g_err[1024] = "";

int fail(int status, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return status;
}

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

bool pdl_allowed() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("ZSV_PDL");
        v = (e && atoi(e) == 0) ? 0 : 1;
    }
    return v != 0;
}

int sm_count() {
    static int n = 0;
    static std::once_flag once;
    std::call_once(once, [] {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return;
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess) n = v;
    });
    return n > 0 ? n : 148;
}

}  // namespace zsv

extern "C" const char* zsv_last_error(void) { return zsv::g_err; }
extern "C" int zsv_abi_version(void) { return 8; }
extern "C" int zsv_cpad(int c) { return zsv::cpad(c); }
extern "C" int zsv_sm_count(void) { return zsv::sm_count(); }
extern "C" unsigned long long zsv_launch_count(void) { return zsv::g_launches.load(std::memory_order_relaxed); }
