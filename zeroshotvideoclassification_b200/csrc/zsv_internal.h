// Host-side helpers shared by the translation units of libzsv_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <utility>

#include "../../include/zsv_b200.h"

namespace zsv {

// Record a message for zsv_last_error() and return the status (printf-style).
int fail(int status, const char* fmt, ...);

inline int cpad(int c) { return (c + 7) & ~7; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// Number of SMs of the current device (cached).
int sm_count();

// Count one kernel launch issued by this library (read back through zsv_launch_count()).
void count_launch();

// Programmatic dependent launch (ZSV_PDL=0 turns it off).  Every kernel of this library is launched with the
// programmatic-stream-serialization attribute and executes griddepcontrol.wait before it touches global memory, so a
// grid is set up and scheduled while the previous kernel of its stream drains (inside a captured CUDA graph the edge
// becomes a programmatic dependency) and results do not change.
bool pdl_allowed();
inline void pdl_attribute(cudaLaunchAttribute* a) {
    a->id = cudaLaunchAttributeProgrammaticStreamSerialization;
    a->val.programmaticStreamSerializationAllowed = pdl_allowed() ? 1 : 0;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
    cudaLaunchAttribute attr;
    pdl_attribute(&attr);
    cfg.attrs = &attr, cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(std::forward<Args>(args))...);
}

// fp32 Linear layers with few batch rows (zsv_linear.cu); shared by the C3D entry points and the embedding head.
// ws may be NULL for linear_forward (the reduction is then not split over blocks).
size_t linear_workspace_bytes(int B, int K, int J);
int linear_forward(const float* x, const float* w, const float* bias, float* out, int B, int K, int J, int relu,
                   float* ws, size_t ws_bytes, cudaStream_t st);
int linear_dgrad(const float* g, const float* w, const float* act_in, float* dx, int B, int K, int J, float* ws,
                 size_t ws_bytes, cudaStream_t st);
int linear_wgrad(const float* g, const float* x, float* dw, float* db, int B, int K, int J, cudaStream_t st);

#define ZSV_CUDA_CHECK(expr)                                                                            \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            return ::zsv::fail(ZSV_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),    \
                               __FILE__, __LINE__);                                                     \
    } while (0)

#define ZSV_LAUNCH_CHECK(name)                                                                          \
    do {                                                                                                \
        cudaError_t _e = cudaGetLastError();                                                            \
        if (_e != cudaSuccess)                                                                          \
            return ::zsv::fail(ZSV_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(_e));  \
        ::zsv::count_launch();                                                                          \
    } while (0)

}  // namespace zsv
