// Embedding head and zero-shot nearest-class search.
//
// Head (network.py:595-596, MLP network.py:603-618): mean over (T,H,W) -> Linear -> ReLU -> Linear ->
// F.normalize, forward and analytic backward, all fp32 on CUDA cores: with B ~ 22 rows these are
// launch/latency-bound weight-streaming kernels, not tensor-core work.
// Loss (main.py:130,179): MSELoss(mean).
// Nearest class (main.py:183, main.py:321-322): scipy cdist(...,'cosine') + argmin / argsort[:, :k],
// restated in fp64 with scipy's exact operation order so that indices are bit-identical.
#include <algorithm>
#include <string.h>
#include <mutex>

#include "zsv_internal.h"
#include "zsv_ptx.cuh"

namespace zsv {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// pooled[b][c] = mean_p feat[b][p][c]
__global__ void pool_kernel(const __nv_bfloat16* __restrict__ feat, float* __restrict__ pooled, int B, int P, int C,
                            int Cp) {
    const int b = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float acc = 0.f;
    for (int p = 0; p < P; ++p) acc += __bfloat162float(feat[((long long)b * P + p) * Cp + c]);
    pooled[(long long)b * C + c] = acc / (float)P;
}

// emb[b] = o[b] / max(||o[b]||, eps); one warp per row
__global__ void normalize_fwd_kernel(const float* __restrict__ o, float* __restrict__ emb, float* __restrict__ onorm,
                                     int B, int E, float eps) {
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    float ss = 0.f;
    for (int e = lane; e < E; e += 32) {
        const float v = o[(long long)b * E + e];
        ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    const float nrm = sqrtf(ss);
    const float d = fmaxf(nrm, eps);
    for (int e = lane; e < E; e += 32) emb[(long long)b * E + e] = o[(long long)b * E + e] / d;
    if (lane == 0) onorm[b] = nrm;
}

// do[b] = (demb - emb * <emb, demb>) / ||o||   (or demb / eps when the norm was clamped)
__global__ void normalize_bwd_kernel(const float* __restrict__ demb, const float* __restrict__ emb,
                                     const float* __restrict__ onorm, float* __restrict__ dout, int B, int E,
                                     float eps) {
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    const float nrm = onorm[b];
    float dot = 0.f;
    for (int e = lane; e < E; e += 32) dot = fmaf(emb[(long long)b * E + e], demb[(long long)b * E + e], dot);
    dot = warp_sum(dot);
    for (int e = lane; e < E; e += 32) {
        const float g = demb[(long long)b * E + e];
        dout[(long long)b * E + e] = nrm > eps ? (g - emb[(long long)b * E + e] * dot) / nrm : g / eps;
    }
}

// dfeat[b][p][c] = dpooled[b][c] / P
__global__ void pool_bwd_kernel(const float* __restrict__ dpooled, __nv_bfloat16* __restrict__ dfeat, int B, int P,
                                int C, int Cp) {
    const long long total = (long long)B * P * Cp;
    const float invP = 1.f / (float)P;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = static_cast<int>(i % Cp);
        const int b = static_cast<int>(i / ((long long)P * Cp));
        dfeat[i] = __float2bfloat16(c < C ? dpooled[(long long)b * C + c] * invP : 0.f);
    }
}

// loss = mean((emb-target)^2) ; demb = 2*(emb-target)/(B*E)*grad_scale ; single block (B*E is tiny)
__global__ void mse_kernel(const float* __restrict__ emb, const float* __restrict__ target, int n, float grad_scale,
                           float* __restrict__ loss, float* __restrict__ demb) {
    __shared__ float sh[32];
    float acc = 0.f;
    const float k = 2.f * grad_scale / (float)n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float d = emb[i] - target[i];
        acc = fmaf(d, d, acc);
        if (demb) demb[i] = k * d;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0 && loss) loss[0] = v / (float)n;
    }
}

// ------------------------------------------------------------------------------------------------
// nearest class.  scipy's cdist_cosine: norms = sqrt(sum of squares), dot = sum of products, both fp64 in the
// order scipy 1.18.1 compiles them (2-lane reduction: even-k and odd-k accumulators over the first D - D%2
// terms, total = even + odd, then the odd tail term; see oracle/nearest_oracle.py).  Products of two fp32
// values are exact in fp64, so fma == mul+add here.  cosine = dot / (nu*nv), clipped to +-1, d = 1 - cosine.
// block = nt threads (one class each per tile) x R embedding rows; distances staged in shared memory,
// then one warp per row extracts the k smallest (lowest index wins ties, NaN sorts last).
// ------------------------------------------------------------------------------------------------
constexpr int kKC = 32;  // k-chunk of the class tile staged in shared memory

template <int R>
__global__ void __launch_bounds__(R == 1 ? 1024 : 128)
nearest_kernel(const float* __restrict__ emb, const float* __restrict__ cls, int N, int C, int D, int k,
               int64_t* __restrict__ idx_out, double* __restrict__ dist_out) {
    extern __shared__ double smd[];
    double* dist = smd;                                               // [R][C]
    double* nrm_e = dist + (size_t)R * C;                         // [R]
    float* s_e = reinterpret_cast<float*>(nrm_e + R);             // [R][D]
    float* s_c = s_e + (size_t)R * D;                             // [nt][kKC+1]
    const int nt = blockDim.x;                                    // classes per tile = threads per block
    const int i0 = blockIdx.x * R;
    const int tid = threadIdx.x;
    for (int i = tid; i < R * D; i += nt) {
        const int r = i / D, kk = i - r * D;
        s_e[i] = (i0 + r < N) ? emb[(long long)(i0 + r) * D + kk] : 0.f;
    }
    __syncthreads();
    const int Deven = D & ~1;
    if (tid < R) {  // scipy _row_norms, same even/odd order as the dot products
        double s0 = 0.0, s1 = 0.0;
        for (int kk = 0; kk < Deven; kk += 2) {
            const double v0 = (double)s_e[tid * D + kk], v1 = (double)s_e[tid * D + kk + 1];
            s0 = __dadd_rn(s0, __dmul_rn(v0, v0));
            s1 = __dadd_rn(s1, __dmul_rn(v1, v1));
        }
        double s = __dadd_rn(s0, s1);
        if (Deven < D) {
            const double v = (double)s_e[tid * D + Deven];
            s = __dadd_rn(s, __dmul_rn(v, v));
        }
        nrm_e[tid] = sqrt(s);
    }
    __syncthreads();
    for (int j0 = 0; j0 < C; j0 += nt) {
        const int j = j0 + tid;
        double acc0[R], acc1[R];  // even-k / odd-k accumulators
        double cc0 = 0.0, cc1 = 0.0;      // squared norm of this thread's class row, same order
#pragma unroll
        for (int r = 0; r < R; ++r) acc0[r] = acc1[r] = 0.0;
        for (int k0 = 0; k0 < D; k0 += kKC) {
            __syncthreads();
            for (int i = tid; i < nt * kKC; i += nt) {
                const int jj = i / kKC, kk = i - jj * kKC;
                s_c[jj * (kKC + 1) + kk] = (j0 + jj < C && k0 + kk < D) ? cls[(long long)(j0 + jj) * D + k0 + kk] : 0.f;
            }
            __syncthreads();
            const int kmax = min(kKC, Deven - k0);  // kKC is even, so chunk-local parity == global parity
            for (int kk = 0; kk < kmax; kk += 2) {
                const double c0 = (double)s_c[tid * (kKC + 1) + kk], c1 = (double)s_c[tid * (kKC + 1) + kk + 1];
                cc0 = __dadd_rn(cc0, __dmul_rn(c0, c0));
                cc1 = __dadd_rn(cc1, __dmul_rn(c1, c1));
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    acc0[r] = __dadd_rn(acc0[r], __dmul_rn((double)s_e[r * D + k0 + kk], c0));
                    acc1[r] = __dadd_rn(acc1[r], __dmul_rn((double)s_e[r * D + k0 + kk + 1], c1));
                }
            }
            if (Deven < D && k0 <= Deven && Deven < k0 + kKC) {  // odd tail term, added after even + odd
                const double ct = (double)s_c[tid * (kKC + 1) + (Deven - k0)];
                cc0 = __dadd_rn(__dadd_rn(cc0, cc1), __dmul_rn(ct, ct));
                cc1 = 0.0;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    acc0[r] = __dadd_rn(__dadd_rn(acc0[r], acc1[r]), __dmul_rn((double)s_e[r * D + Deven], ct));
                    acc1[r] = 0.0;
                }
            }
        }
        if (j < C) {
            // when D is odd the tail step above already folded the lanes (and zeroed the odd one: x + 0.0 == x)
            const double nv = sqrt(__dadd_rn(cc0, cc1));
#pragma unroll
            for (int r = 0; r < R; ++r) {
                double cosine = __dadd_rn(acc0[r], acc1[r]) / (nrm_e[r] * nv);
                if (fabs(cosine) > 1.0) cosine = copysign(1.0, cosine);
                dist[(size_t)r * C + j] = 1.0 - cosine;
            }
        }
    }
    __syncthreads();
    // selection: warp w handles rows w, w + #warps, ...
    const int warp = tid >> 5, lane = tid & 31;
    for (int r = warp; r < R; r += (nt >> 5)) {
        if (i0 + r >= N) continue;
        double* dr = dist + (size_t)r * C;
        for (int sel = 0; sel < k; ++sel) {
            double best = 0.0;
            int bidx = -1;   // -1: nothing yet
            bool bnan = true;
            for (int j = lane; j < C; j += 32) {
                const double v = dr[j];
                if (v == -1.0e300) continue;  // already taken (distances live in [0,2])
                const bool vnan = v != v;
                bool better;
                if (bidx < 0) better = true;
                else if (bnan != vnan) better = bnan;          // a number beats NaN
                else if (vnan) better = false;                 // both NaN: lower index (seen first) stays
                else better = v < best;
                if (better) best = v, bidx = j, bnan = vnan;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
                const int on = __shfl_xor_sync(0xffffffffu, (int)bnan, o);
                bool better;
                if (oi < 0) better = false;
                else if (bidx < 0) better = true;
                else if ((bool)on != bnan) better = bnan;
                else if (bnan) better = oi < bidx;
                else better = (ov < best) || (ov == best && oi < bidx);
                if (better) best = ov, bidx = oi, bnan = (bool)on;
            }
            if (lane == 0) {
                idx_out[(long long)(i0 + r) * k + sel] = bidx;
                if (dist_out) dist_out[(long long)(i0 + r) * k + sel] = best;
            }
            __syncwarp();
            if (bidx >= 0 && (bidx & 31) == lane) dr[bidx] = -1.0e300;
            __syncwarp();
        }
    }
}

}  // namespace
}  // namespace zsv

using namespace zsv;

extern "C" int zsv_head_fwd(const void* feat, int B, int P, int C, const float* w1, const float* b1, int Hd,
                            const float* w2, const float* b2, int E, float eps, float* pooled, float* hidden,
                            float* onorm, float* emb, void* stream) {
    if (!feat || !w1 || !w2 || !pooled || !hidden || !onorm || !emb) return fail(ZSV_ERR_BAD_ARG, "head_fwd: null pointer");
    if (B < 1 || P < 1 || C < 1 || Hd < 1 || E < 1) return fail(ZSV_ERR_BAD_ARG, "head_fwd: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    pool_kernel<<<dim3(ceil_div(C, 128), B), 128, 0, st>>>((const __nv_bfloat16*)feat, pooled, B, P, C, cpad(C));
    ZSV_LAUNCH_CHECK("pool_kernel");
    int rc = linear_forward(pooled, w1, b1, hidden, B, C, Hd, 1, nullptr, 0, st);
    if (rc) return rc;
    // raw projection goes to emb, then normalised in place
    rc = linear_forward(hidden, w2, b2, emb, B, Hd, E, 0, nullptr, 0, st);
    if (rc) return rc;
    normalize_fwd_kernel<<<ceil_div(B * 32, 128), 128, 0, st>>>(emb, emb, onorm, B, E, eps);
    ZSV_LAUNCH_CHECK("normalize_fwd_kernel");
    return ZSV_OK;
}

extern "C" size_t zsv_head_bwd_scratch(int B, int C, int Hd, int E) {
    if (B < 1 || C < 1 || Hd < 1 || E < 1) return 0;
    const size_t fixed = sizeof(float) * (size_t)B * ((size_t)E + Hd + C);
    return fixed + std::max(linear_workspace_bytes(B, Hd, E), linear_workspace_bytes(B, C, Hd));
}

extern "C" int zsv_head_bwd(const float* demb, const float* emb, const float* onorm, const float* pooled,
                            const float* hidden, int B, int P, int C, const float* w1, int Hd, const float* w2, int E,
                            float eps, float* dw1, float* db1, float* dw2, float* db2, void* dfeat, float* scratch,
                            size_t scratch_bytes, void* stream) {
    if (!demb || !emb || !onorm || !pooled || !hidden || !w1 || !w2 || !scratch)
        return fail(ZSV_ERR_BAD_ARG, "head_bwd: null pointer");
    if (scratch_bytes < zsv_head_bwd_scratch(B, C, Hd, E))
        return fail(ZSV_ERR_WORKSPACE, "head_bwd: scratch %zu < required %zu bytes", scratch_bytes, zsv_head_bwd_scratch(B, C, Hd, E));
    cudaStream_t st = (cudaStream_t)stream;
    // scratch: do [B][E] | dh [B][Hd] | dpooled [B][C] | split partials of the data gradients
    float* dout = scratch;
    float* dh = dout + (size_t)B * E;
    float* dpooled = dh + (size_t)B * Hd;
    float* ws = dpooled + (size_t)B * C;
    const size_t ws_bytes = scratch_bytes - sizeof(float) * (size_t)B * ((size_t)E + Hd + C);
    normalize_bwd_kernel<<<ceil_div(B * 32, 128), 128, 0, st>>>(demb, emb, onorm, dout, B, E, eps);
    ZSV_LAUNCH_CHECK("normalize_bwd_kernel");
    int rc;
    if (dw2) {
        rc = linear_wgrad(dout, hidden, dw2, db2, B, Hd, E, st);
        if (rc) return rc;
    }
    rc = linear_dgrad(dout, w2, hidden, dh, B, Hd, E, ws, ws_bytes, st);      // ReLU mask of the hidden layer applied to dh
    if (rc) return rc;
    if (dw1) {
        rc = linear_wgrad(dh, pooled, dw1, db1, B, C, Hd, st);
        if (rc) return rc;
    }
    if (dfeat) {
        rc = linear_dgrad(dh, w1, nullptr, dpooled, B, C, Hd, ws, ws_bytes, st);
        if (rc) return rc;
        const long long total = (long long)B * P * cpad(C);
        pool_bwd_kernel<<<(int)std::min<long long>(ceil_div_ll(total, 256), 148 * 8), 256, 0, st>>>(
            dpooled, (__nv_bfloat16*)dfeat, B, P, C, cpad(C));
        ZSV_LAUNCH_CHECK("pool_bwd_kernel");
    }
    return ZSV_OK;
}

extern "C" int zsv_l2norm_fwd(const float* o, float* emb, float* onorm, int B, int E, float eps, void* stream) {
    if (!o || !emb || !onorm) return fail(ZSV_ERR_BAD_ARG, "l2norm_fwd: null pointer");
    normalize_fwd_kernel<<<ceil_div(B * 32, 128), 128, 0, (cudaStream_t)stream>>>(o, emb, onorm, B, E, eps);
    ZSV_LAUNCH_CHECK("normalize_fwd_kernel");
    return ZSV_OK;
}

extern "C" int zsv_l2norm_bwd(const float* demb, const float* emb, const float* onorm, float* dout, int B, int E,
                              float eps, void* stream) {
    if (!demb || !emb || !onorm || !dout) return fail(ZSV_ERR_BAD_ARG, "l2norm_bwd: null pointer");
    normalize_bwd_kernel<<<ceil_div(B * 32, 128), 128, 0, (cudaStream_t)stream>>>(demb, emb, onorm, dout, B, E, eps);
    ZSV_LAUNCH_CHECK("normalize_bwd_kernel");
    return ZSV_OK;
}

extern "C" int zsv_mse_fwd_bwd(const float* emb, const float* target, int B, int E, float grad_scale, float* loss,
                               float* demb, void* stream) {
    if (!emb || !target) return fail(ZSV_ERR_BAD_ARG, "mse: null pointer");
    mse_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(emb, target, B * E, grad_scale, loss, demb);
    ZSV_LAUNCH_CHECK("mse_kernel");
    return ZSV_OK;
}

extern "C" int zsv_nearest_class(const float* emb, const float* cls, int N, int C, int D, int k, int64_t* idx_out,
                                 double* dist_out, void* stream) {
    if (!emb || !cls || !idx_out) return fail(ZSV_ERR_BAD_ARG, "nearest_class: null pointer");
    if (N < 0 || C < 1 || D < 1 || k < 1 || k > 8 || k > C) return fail(ZSV_ERR_BAD_ARG, "nearest_class: bad sizes");
    if (N == 0) return ZSV_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // few rows (train-time batch, main.py:183): one row per block so the grid still covers the SMs;
    // many rows (evaluation, main.py:321): 8 rows per block amortise the class-table traffic
    const int R = N <= 2048 ? 1 : 8;
    // R == 1: the fp64 chain of one (row, class) pair is sequential by construction (scipy's order), so the block is made
    // as wide as the class table (up to 1024 threads) instead of walking 128-class tiles one after the other
    const int nt = R == 1 ? std::min(1024, ceil_div(C, 128) * 128) : 128;
    const size_t smem = sizeof(double) * ((size_t)R * C + R) + sizeof(float) * ((size_t)R * D + (size_t)nt * (kKC + 1));
    if (smem > 200 * 1024)
        return fail(ZSV_ERR_UNSUPPORTED, "nearest_class: class table too large for shared memory (%zu bytes)", smem);
    static std::once_flag attr_once;
    std::call_once(attr_once, [] {
        cudaFuncSetAttribute(nearest_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(nearest_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    });
    if (R == 1)
        nearest_kernel<1><<<N, nt, smem, st>>>(emb, cls, N, C, D, k, idx_out, dist_out);
    else
        nearest_kernel<8><<<ceil_div(N, 8), 128, smem, st>>>(emb, cls, N, C, D, k, idx_out, dist_out);
    ZSV_LAUNCH_CHECK("nearest_kernel");
    return ZSV_OK;
}
