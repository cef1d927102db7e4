// fp32 Linear layers with a handful of batch rows (C3D fc6 8192->4096 and regressor 4096->300, network.py:120,132,166,178;
// the MLP head 512->512->300, network.py:603-618): forward, data gradient and weight gradient.
//
// With B ~ 22 rows these are weight-STREAMING kernels: every pass moves the weight matrix through the SMs exactly once
// (fc6: 134 MB, 20 us at HBM rate) and spends B FMAs per weight on CUDA cores (fp32, exact products; the tensor cores
// would buy nothing on an HBM-bound pass).  The previous kernels kept one weight element per thread per load in flight
// (fc6: forward 744 us, dgrad 6.3 ms, wgrad 380 us -- half of the C3D training step); here every thread keeps several
// 16-byte weight loads in flight and reuses each activation value for 4 weight rows (forward) / 4 columns (dgrad,
// wgrad) from registers, so shared-memory reads stay below the FMA rate.
//   forward : warp = 4 output features, lanes stride over k with float4 loads, x staged in shared memory per k chunk;
//             the k range is split over blockIdx.y when the matrix has few rows (deterministic two-pass combine)
//   dgrad   : thread = 4 consecutive k for all batch rows, loop over output features j with coalesced weight rows,
//             j range split over blockIdx.y, partials combined (and ReLU-masked) by a second kernel
//   wgrad   : thread = 4 consecutive k, x rows held in registers, one 16-byte store per (j, k4): write-bound
// Batches larger than kBT rows are processed in passes of kBT rows (the weights are streamed once per pass).
#include <algorithm>
#include <mutex>
#include <string.h>

#include "zsv_internal.h"
#include "zsv_ptx.cuh"

namespace zsv {
namespace {

constexpr int kBT = 24;        // batch rows per pass (accumulators per thread: kBT x 4)
constexpr int kKC = 512;       // forward: k chunk of x staged in shared memory
constexpr int kFwdRows = 32;   // forward: output features per block (8 warps x 4)
constexpr int kDgK = 1024;     // dgrad / wgrad: k columns per block (256 threads x 4)
constexpr int kDgJ = 64;       // dgrad: output features staged per shared-memory pass
constexpr long long kSmallMatrix = 1LL << 20;   // weights up to here: the latency-bound kernels for small matrices

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// out[b][j] (or partial[split][b][j]) = sum_{k in split} x[b][k] * w[j][k]
// grid = (ceil(J/32), splits, batch passes); K % 4 == 0 and 16-byte aligned rows (checked by the host), else scalar path
template <bool kVec>
__global__ void __launch_bounds__(256)
linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                  float* __restrict__ out, float* __restrict__ partial, int B, int K, int J, int relu, int k_per_split) {
    pdl_wait();
    extern __shared__ float xs[];                       // [kBT][kKC + 4]
    constexpr int kPitch = kKC + 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j0 = blockIdx.x * kFwdRows + warp * 4;
    const int b0 = blockIdx.z * kBT;
    const int nb = min(kBT, B - b0);
    const int kbeg = blockIdx.y * k_per_split, kend = min(K, kbeg + k_per_split);
    float acc[kBT][4];
#pragma unroll
    for (int b = 0; b < kBT; ++b)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[b][r] = 0.f;
    const float* wr[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) wr[r] = w + (long long)min(j0 + r, J - 1) * K;   // rows past J are computed and dropped
    for (int k0 = kbeg; k0 < kend; k0 += kKC) {
        const int kc = min(kKC, kend - k0);
        __syncthreads();
        for (int i = threadIdx.x; i < kBT * (kKC / 4); i += 256) {      // stage x[b0.., k0..k0+kc) (zero padded)
            const int b = i / (kKC / 4), q = (i - b * (kKC / 4)) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (b < nb) {
                const float* src = x + (long long)(b0 + b) * K + k0 + q;
                if (kVec && q + 3 < kc) v = *reinterpret_cast<const float4*>(src);
                else {
                    if (q < kc) v.x = src[0];
                    if (q + 1 < kc) v.y = src[1];
                    if (q + 2 < kc) v.z = src[2];
                    if (q + 3 < kc) v.w = src[3];
                }
            }
            *reinterpret_cast<float4*>(xs + b * kPitch + q) = v;
        }
        __syncthreads();
#pragma unroll 1
        for (int q = lane * 4; q < kc; q += 128) {
            float4 wv[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float* src = wr[r] + k0 + q;
                if (kVec && q + 3 < kc) wv[r] = *reinterpret_cast<const float4*>(src);
                else {
                    wv[r] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (q < kc) wv[r].x = src[0];
                    if (q + 1 < kc) wv[r].y = src[1];
                    if (q + 2 < kc) wv[r].z = src[2];
                    if (q + 3 < kc) wv[r].w = src[3];
                }
            }
#pragma unroll
            for (int b = 0; b < kBT; ++b) {
                const float4 xv = *reinterpret_cast<const float4*>(xs + b * kPitch + q);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    acc[b][r] = fmaf(wv[r].x, xv.x, acc[b][r]);
                    acc[b][r] = fmaf(wv[r].y, xv.y, acc[b][r]);
                    acc[b][r] = fmaf(wv[r].z, xv.z, acc[b][r]);
                    acc[b][r] = fmaf(wv[r].w, xv.w, acc[b][r]);
                }
            }
        }
    }
#pragma unroll
    for (int b = 0; b < kBT; ++b)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float s = warp_sum(acc[b][r]);
            if (lane == 0 && b < nb && j0 + r < J) {
                if (partial != nullptr) {
                    partial[((long long)blockIdx.y * B + b0 + b) * J + j0 + r] = s;
                } else {
                    float v = s + (bias ? bias[j0 + r] : 0.f);
                    if (relu) v = fmaxf(v, 0.f);
                    out[(long long)(b0 + b) * J + j0 + r] = v;
                }
            }
        }
}

// out[b][j] = act(sum_s partial[s][b][j] + bias[j])   (fixed order: deterministic)
__global__ void linear_fwd_finish_kernel(const float* __restrict__ partial, const float* __restrict__ bias,
                                         float* __restrict__ out, int splits, long long n, int J, int relu) {
    pdl_wait();
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int sp = 0; sp < splits; ++sp) s += partial[sp * n + i];
    s += bias ? bias[i % J] : 0.f;
    out[i] = relu ? fmaxf(s, 0.f) : s;
}

// dst[split][b][k] (or dx[b][k] when gridDim.y == 1 and no mask) = sum_{j in split} g[b][j] * w[j][k]
// grid = (ceil(K/1024), splits, batch passes)
template <bool kVec>
__global__ void __launch_bounds__(256)
linear_dgrad_kernel(const float* __restrict__ g, const float* __restrict__ w, float* __restrict__ dst, int B, int K, int J,
                    int j_per_split) {
    pdl_wait();
    __shared__ __align__(16) float gs[kDgJ][kBT];       // g[b][j] transposed: one 16-byte broadcast load = 4 batch rows
    const int k = blockIdx.x * kDgK + threadIdx.x * 4;
    const int b0 = blockIdx.z * kBT;
    const int nb = min(kBT, B - b0);
    const int jbeg = blockIdx.y * j_per_split, jend = min(J, jbeg + j_per_split);
    float acc[kBT][4];
#pragma unroll
    for (int b = 0; b < kBT; ++b)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[b][r] = 0.f;
    for (int j0 = jbeg; j0 < jend; j0 += kDgJ) {
        const int jc = min(kDgJ, jend - j0);
        __syncthreads();
        for (int i = threadIdx.x; i < kDgJ * kBT; i += 256) {
            const int jj = i / kBT, b = i - jj * kBT;
            gs[jj][b] = (jj < jc && b < nb) ? g[(long long)(b0 + b) * J + j0 + jj] : 0.f;
        }
        __syncthreads();
        if (k < K) {
#pragma unroll 4
            for (int jj = 0; jj < jc; ++jj) {
                const float* src = w + (long long)(j0 + jj) * K + k;
                float4 wv;
                if (kVec) wv = *reinterpret_cast<const float4*>(src);
                else {
                    wv.x = src[0];
                    wv.y = k + 1 < K ? src[1] : 0.f;
                    wv.z = k + 2 < K ? src[2] : 0.f;
                    wv.w = k + 3 < K ? src[3] : 0.f;
                }
#pragma unroll
                for (int b4 = 0; b4 < kBT / 4; ++b4) {
                    const float4 gv = *reinterpret_cast<const float4*>(&gs[jj][b4 * 4]);
                    const float gb[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        acc[b4 * 4 + t][0] = fmaf(gb[t], wv.x, acc[b4 * 4 + t][0]);
                        acc[b4 * 4 + t][1] = fmaf(gb[t], wv.y, acc[b4 * 4 + t][1]);
                        acc[b4 * 4 + t][2] = fmaf(gb[t], wv.z, acc[b4 * 4 + t][2]);
                        acc[b4 * 4 + t][3] = fmaf(gb[t], wv.w, acc[b4 * 4 + t][3]);
                    }
                }
            }
        }
    }
    if (k >= K) return;
#pragma unroll
    for (int b = 0; b < kBT; ++b) {
        if (b < nb) {
            float* d = dst + ((long long)blockIdx.y * B + b0 + b) * K + k;
            if (kVec) *reinterpret_cast<float4*>(d) = make_float4(acc[b][0], acc[b][1], acc[b][2], acc[b][3]);
            else {
                d[0] = acc[b][0];
                if (k + 1 < K) d[1] = acc[b][1];
                if (k + 2 < K) d[2] = acc[b][2];
                if (k + 3 < K) d[3] = acc[b][3];
            }
        }
    }
}

// dx[i] = (sum_s partial[s][i]) * (act ? [act[i] > 0] : 1)
__global__ void linear_dgrad_finish_kernel(const float* __restrict__ partial, const float* __restrict__ act,
                                           float* __restrict__ dx, int splits, long long n) {
    pdl_wait();
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int sp = 0; sp < splits; ++sp) s += partial[sp * n + i];
    dx[i] = (act && !(act[i] > 0.f)) ? 0.f : s;
}

// dw[j][k] = sum_b g[b][j] * x[b][k];  db[j] = sum_b g[b][j].   grid = (ceil(K/1024), ceil(J/jt)); x rows in registers.
template <bool kVec>
__global__ void __launch_bounds__(256)
linear_wgrad_kernel(const float* __restrict__ g, const float* __restrict__ x, float* __restrict__ dw,
                    float* __restrict__ db, int B, int K, int J, int jt) {
    pdl_wait();
    extern __shared__ float gsm[];                      // [jt][kBT] per batch pass
    const int k = blockIdx.x * kDgK + threadIdx.x * 4;
    const int j0 = blockIdx.y * jt;
    const int jc = min(jt, J - j0);
    const int passes = (B + kBT - 1) / kBT;
    for (int ps = 0; ps < passes; ++ps) {
        const int b0 = ps * kBT;
        const int nb = min(kBT, B - b0);
        __syncthreads();
        for (int i = threadIdx.x; i < jt * kBT; i += 256) {
            const int jj = i / kBT, b = i - jj * kBT;
            gsm[i] = (jj < jc && b < nb) ? g[(long long)(b0 + b) * J + j0 + jj] : 0.f;
        }
        __syncthreads();
        if (blockIdx.x == 0 && db != nullptr)
            for (int jj = threadIdx.x; jj < jc; jj += 256) {
                float s = ps == 0 ? 0.f : db[j0 + jj];
                for (int b = 0; b < kBT; ++b) s += gsm[jj * kBT + b];
                db[j0 + jj] = s;
            }
        if (k >= K) continue;
        float4 xv[kBT];
#pragma unroll
        for (int b = 0; b < kBT; ++b) {
            xv[b] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (b < nb) {
                const float* src = x + (long long)(b0 + b) * K + k;
                if (kVec) xv[b] = *reinterpret_cast<const float4*>(src);
                else {
                    xv[b].x = src[0];
                    if (k + 1 < K) xv[b].y = src[1];
                    if (k + 2 < K) xv[b].z = src[2];
                    if (k + 3 < K) xv[b].w = src[3];
                }
            }
        }
#pragma unroll 2
        for (int jj = 0; jj < jc; ++jj) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int b4 = 0; b4 < kBT / 4; ++b4) {
                const float4 gv = *reinterpret_cast<const float4*>(gsm + jj * kBT + b4 * 4);
                const float gb[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    o.x = fmaf(gb[t], xv[b4 * 4 + t].x, o.x);
                    o.y = fmaf(gb[t], xv[b4 * 4 + t].y, o.y);
                    o.z = fmaf(gb[t], xv[b4 * 4 + t].z, o.z);
                    o.w = fmaf(gb[t], xv[b4 * 4 + t].w, o.w);
                }
            }
            float* d = dw + (long long)(j0 + jj) * K + k;
            if (ps > 0) {   // later batch passes accumulate onto the first (same thread, same address: ordered)
                if (kVec) {
                    const float4 p = *reinterpret_cast<const float4*>(d);
                    o.x += p.x, o.y += p.y, o.z += p.z, o.w += p.w;
                } else {
                    o.x += d[0];
                    if (k + 1 < K) o.y += d[1];
                    if (k + 2 < K) o.z += d[2];
                    if (k + 3 < K) o.w += d[3];
                }
            }
            if (kVec) *reinterpret_cast<float4*>(d) = o;
            else {
                d[0] = o.x;
                if (k + 1 < K) d[1] = o.y;
                if (k + 2 < K) d[2] = o.z;
                if (k + 3 < K) d[3] = o.w;
            }
        }
    }
}

// ---- small weight matrices (the MLP head: 512 x 512, 300 x 512): latency-bound, parallelism over (row, output) -------
// out[b][j] = act(sum_k x[b][k] * w[j][k] + bias[j]); one warp per output feature j, 8 batch rows per block row
__global__ void linear_fwd_small_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                        const float* __restrict__ bias, float* __restrict__ out, int B, int K, int J,
                                        int relu) {
    pdl_wait();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= J) return;
    const float* wr = w + (long long)warp * K;
    for (int b0 = blockIdx.y * 8; b0 < B; b0 += 8 * gridDim.y) {
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        for (int k = lane; k < K; k += 32) {
            const float wv = wr[k];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (b0 + i < B) acc[i] = fmaf(wv, x[(long long)(b0 + i) * K + k], acc[i]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float s = warp_sum(acc[i]);
            if (lane == 0 && b0 + i < B) {
                float v = s + (bias ? bias[warp] : 0.f);
                if (relu) v = fmaxf(v, 0.f);
                out[(long long)(b0 + i) * J + warp] = v;
            }
        }
    }
}

// dx[b][k] = (sum_j g[b][j] * w[j][k]) * (act ? [act[b][k] > 0] : 1): block = 32 k lanes x 8 slices of the j reduction,
// 8 batch rows; the slices are combined in fixed order through shared memory.  grid = (ceil(K/32), ceil(B/8)).
__global__ void __launch_bounds__(256)
linear_dgrad_small_kernel(const float* __restrict__ g, const float* __restrict__ w, const float* __restrict__ act,
                          float* __restrict__ dx, int B, int K, int J) {
    pdl_wait();
    __shared__ float red[8][8][33];
    const int kl = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int k = blockIdx.x * 32 + kl;
    const int b0 = blockIdx.y * 8;
    const int jper = (J + 7) >> 3;
    const int j0 = sl * jper, j1 = min(J, j0 + jper);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll 4
    for (int j = j0; j < j1; ++j) {
        const float wv = k < K ? w[(long long)j * K + k] : 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (b0 + i < B) acc[i] = fmaf(g[(long long)(b0 + i) * J + j], wv, acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) red[sl][i][kl] = acc[i];
    __syncthreads();
    const int i = sl;   // 8 warps -> 8 batch rows
    if (b0 + i < B && k < K) {
        float s2 = 0.f;
#pragma unroll
        for (int t = 0; t < 8; ++t) s2 += red[t][i][kl];
        const long long o = (long long)(b0 + i) * K + k;
        dx[o] = (act && !(act[o] > 0.f)) ? 0.f : s2;
    }
}

__global__ void relu_mask_kernel(const float* __restrict__ dy, const float* __restrict__ act, float* __restrict__ out,
                                 long long n) {
    pdl_wait();
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = act[i] > 0.f ? dy[i] : 0.f;
}

bool vec_ok(const void* p, int K) { return (K & 3) == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// how the k (forward) / j (dgrad) reduction is split over blockIdx.y so that about two blocks per SM stream the weights
int fwd_splits(int K, int J) {
    const int jb = ceil_div(J, kFwdRows);
    return std::max(1, std::min(ceil_div(K, kKC), (2 * sm_count()) / jb));
}
int dgrad_splits(int K, int J) {
    const int kb = ceil_div(K, kDgK);
    return std::max(1, std::min(ceil_div(J, kDgJ), (2 * sm_count()) / kb));
}

}  // namespace

size_t linear_workspace_bytes(int B, int K, int J) {
    const size_t f = (size_t)fwd_splits(K, J) * B * J, d = (size_t)dgrad_splits(K, J) * B * K;
    return sizeof(float) * (std::max(f, d) + (size_t)B * J);   // split partials + the ReLU-masked gradient
}

int linear_forward(const float* x, const float* w, const float* bias, float* out, int B, int K, int J, int relu,
                   float* ws, size_t ws_bytes, cudaStream_t st) {
    if ((long long)K * J <= kSmallMatrix) {
        zsv::launch(linear_fwd_small_kernel, dim3(ceil_div(J * 32, 256), ceil_div(B, 8)), 256, 0, st, x, w, bias, out, B, K, J, relu);
        ZSV_LAUNCH_CHECK("linear_fwd_small_kernel");
        return ZSV_OK;
    }
    const int splits = ws ? fwd_splits(K, J) : 1;       // no workspace: the k reduction stays inside one block
    if (splits > 1 && ws_bytes < sizeof(float) * (size_t)splits * B * J)
        return fail(ZSV_ERR_WORKSPACE, "linear_fwd: workspace too small (%zu bytes)", ws_bytes);
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    constexpr int smem = kBT * (kKC + 4) * (int)sizeof(float);
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(linear_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (attr_err == cudaSuccess)
            attr_err = cudaFuncSetAttribute(linear_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    });
    if (attr_err != cudaSuccess) return fail(ZSV_ERR_CUDA, "cudaFuncSetAttribute(linear_fwd) failed: %s", cudaGetErrorString(attr_err));
    const int kps = ceil_div(ceil_div(K, splits), kKC) * kKC;   // whole chunks per split
    const int nsplit = ceil_div(K, kps);
    const dim3 grid(ceil_div(J, kFwdRows), nsplit, ceil_div(B, kBT));
    float* partial = nsplit > 1 ? ws : nullptr;
    if (vec_ok(x, K) && vec_ok(w, K))
        zsv::launch(linear_fwd_kernel<true>, grid, 256, smem, st, x, w, bias, out, partial, B, K, J, relu, kps);
    else
        zsv::launch(linear_fwd_kernel<false>, grid, 256, smem, st, x, w, bias, out, partial, B, K, J, relu, kps);
    ZSV_LAUNCH_CHECK("linear_fwd_kernel");
    if (nsplit > 1) {
        const long long n = (long long)B * J;
        zsv::launch(linear_fwd_finish_kernel, (int)ceil_div_ll(n, 256), 256, 0, st, partial, bias, out, nsplit, n, J, relu);
        ZSV_LAUNCH_CHECK("linear_fwd_finish_kernel");
    }
    return ZSV_OK;
}

// g: gradient w.r.t. the layer output, ALREADY masked by the caller when the layer had a ReLU; act_in: optional forward
// activation of the layer INPUT whose ReLU mask is applied to dx (the MLP head's hidden layer)
int linear_dgrad(const float* g, const float* w, const float* act_in, float* dx, int B, int K, int J, float* ws,
                 size_t ws_bytes, cudaStream_t st) {
    if ((long long)K * J <= kSmallMatrix) {
        zsv::launch(linear_dgrad_small_kernel, dim3(ceil_div(K, 32), ceil_div(B, 8)), 256, 0, st, g, w, act_in, dx, B, K, J);
        ZSV_LAUNCH_CHECK("linear_dgrad_small_kernel");
        return ZSV_OK;
    }
    const int splits = dgrad_splits(K, J);
    const int jps = ceil_div(ceil_div(J, splits), kDgJ) * kDgJ;
    const int nsplit = ceil_div(J, jps);
    const bool direct = nsplit == 1 && act_in == nullptr;
    if (!direct && (!ws || ws_bytes < sizeof(float) * (size_t)nsplit * B * K))
        return fail(ZSV_ERR_WORKSPACE, "linear_dgrad: workspace too small (%zu bytes)", ws_bytes);
    const dim3 grid(ceil_div(K, kDgK), nsplit, ceil_div(B, kBT));
    float* dst = direct ? dx : ws;
    if (vec_ok(w, K) && vec_ok(dst, K))
        zsv::launch(linear_dgrad_kernel<true>, grid, 256, 0, st, g, w, dst, B, K, J, jps);
    else
        zsv::launch(linear_dgrad_kernel<false>, grid, 256, 0, st, g, w, dst, B, K, J, jps);
    ZSV_LAUNCH_CHECK("linear_dgrad_kernel");
    if (!direct) {
        const long long n = (long long)B * K;
        zsv::launch(linear_dgrad_finish_kernel, (int)ceil_div_ll(n, 256), 256, 0, st, ws, act_in, dx, nsplit, n);
        ZSV_LAUNCH_CHECK("linear_dgrad_finish_kernel");
    }
    return ZSV_OK;
}

int linear_wgrad(const float* g, const float* x, float* dw, float* db, int B, int K, int J, cudaStream_t st) {
    const int kb = ceil_div(K, kDgK);
    // rows per block: enough blocks to cover the SMs twice, at least 8 rows so the x tile in registers is reused
    int jt = 64;
    while (jt > 8 && (long long)kb * ceil_div(J, jt) < 2LL * sm_count()) jt >>= 1;
    const dim3 grid(kb, ceil_div(J, jt));
    const size_t smem = (size_t)jt * kBT * sizeof(float);
    if (vec_ok(x, K) && vec_ok(dw, K))
        zsv::launch(linear_wgrad_kernel<true>, grid, 256, smem, st, g, x, dw, db, B, K, J, jt);
    else
        zsv::launch(linear_wgrad_kernel<false>, grid, 256, smem, st, g, x, dw, db, B, K, J, jt);
    ZSV_LAUNCH_CHECK("linear_wgrad_kernel");
    return ZSV_OK;
}

int relu_mask(const float* dy, const float* act, float* out, long long n, cudaStream_t st) {
    zsv::launch(relu_mask_kernel, (int)ceil_div_ll(n, 256), 256, 0, st, dy, act, out, n);
    ZSV_LAUNCH_CHECK("relu_mask_kernel");
    return ZSV_OK;
}

}  // namespace zsv

using namespace zsv;

extern "C" size_t zsv_linear_workspace(int B, int K, int J) {
    if (B < 1 || K < 1 || J < 1) return 0;
    return linear_workspace_bytes(B, K, J);
}

extern "C" int zsv_linear_fwd(const float* x, const float* w, const float* bias, float* out, int B, int K, int J,
                              int relu, void* workspace, size_t workspace_bytes, void* stream) {
    if (!x || !w || !out) return fail(ZSV_ERR_BAD_ARG, "linear_fwd: null pointer");
    if (B < 1 || K < 1 || J < 1) return fail(ZSV_ERR_BAD_ARG, "linear_fwd: bad sizes");
    return linear_forward(x, w, bias, out, B, K, J, relu, (float*)workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int zsv_linear_bwd(const float* dy, const float* x, const float* w, const float* act, int B, int K, int J,
                              float* dx, float* dw, float* db, void* workspace, size_t workspace_bytes, void* stream) {
    if (!dy || !x || !w) return fail(ZSV_ERR_BAD_ARG, "linear_bwd: null pointer");
    if (B < 1 || K < 1 || J < 1) return fail(ZSV_ERR_BAD_ARG, "linear_bwd: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    float* ws = (float*)workspace;
    const size_t need = linear_workspace_bytes(B, K, J);
    if (!ws || workspace_bytes < need)
        return fail(ZSV_ERR_WORKSPACE, "linear_bwd: workspace %zu < required %zu bytes", workspace_bytes, need);
    const float* g = dy;
    float* masked = ws + (need / sizeof(float) - (size_t)B * J);   // tail of the workspace
    if (act) {  // ReLU on the forward output: mask dy first
        int rc = relu_mask(dy, act, masked, (long long)B * J, st);
        if (rc) return rc;
        g = masked;
    }
    if (dw) {
        int rc = linear_wgrad(g, x, dw, db, B, K, J, st);
        if (rc) return rc;
    }
    if (dx) {
        int rc = linear_dgrad(g, w, nullptr, dx, B, K, J, ws, need - sizeof(float) * (size_t)B * J, st);
        if (rc) return rc;
    }
    return ZSV_OK;
}
