// Error reporting and device queries shared by all entry points of libzsv_b200.so.
#include <mutex>
#include <string.h>

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
extern "C" int zsv_abi_version(void) { return 1; }
extern "C" int zsv_cpad(int c) { return zsv::cpad(c); }
