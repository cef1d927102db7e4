// HBM-bound kernels of the hot path: layout conversion at the PyTorch boundary, BatchNorm3d finalize /
// apply (+ReLU, +residual) / backward, MaxPool3d.  All activations are bf16 NDHWC with channel pitch cpad(C);
// every thread moves 16-byte vectors (8 channels) so that warps read and write full 128-byte lines.
//
// Reference call sites: nn.BatchNorm3d resnet.py:48,95,97,182,185,272; nn.ReLU resnet.py:49,95,98;
// residual add resnet.py:110-111; nn.MaxPool3d network.py:103-118; input reshape network.py:534-535.
#include <algorithm>
#include <mutex>

#include "zsv_internal.h"
#include "zsv_ptx.cuh"

namespace zsv {
namespace {

constexpr int kMaxC = 2048;  // largest channel pitch staged in shared memory by the BN kernels

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
    f[0] = bf16_lo(v.x), f[1] = bf16_hi(v.x), f[2] = bf16_lo(v.y), f[3] = bf16_hi(v.y);
    f[4] = bf16_lo(v.z), f[5] = bf16_hi(v.z), f[6] = bf16_lo(v.w), f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                      pack_bf16x2(f[6], f[7]));
}

// ------------------------------------------------------------------------------------------------
// layout conversion
// ------------------------------------------------------------------------------------------------
// fp32 NCDHW -> bf16 [N][T][H][Wp][Cp]; column w of the source lands at column w + wl; pad columns/lanes are zero.
__global__ void ncdhw_to_ndhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int C,
                                      int T, int H, int W, int Cp, int Wp, int wl) {
    pdl_wait();
    const int V = Cp >> 3;
    const long long thw = (long long)T * H * W;
    const long long total = (long long)N * T * H * Wp * V;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        // position fastest so that the strided fp32 reads of one channel are coalesced across the warp
        long long r = i;
        const int wp = static_cast<int>(r % Wp);
        r /= Wp;
        const int h = static_cast<int>(r % H);
        r /= H;
        const int t = static_cast<int>(r % T);
        r /= T;
        const int n = static_cast<int>(r % N);
        const int g = static_cast<int>(r / N);
        const int w = wp - wl;
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = g * 8 + j;
            f[j] = (c < C && w >= 0 && w < W) ? x[((long long)n * C + c) * thw + ((long long)t * H + h) * W + w] : 0.f;
        }
        const long long pos = (((long long)n * T + t) * H + h) * Wp + wp;
        *reinterpret_cast<uint4*>(out + pos * Cp + g * 8) = pack8(f);
    }
}

__global__ void ndhwc_to_ncdhw_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int N, int C,
                                      int T, int H, int W, int Cp) {
    pdl_wait();
    const int V = Cp >> 3;
    const long long thw = (long long)T * H * W;
    const long long total = (long long)N * thw * V;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        const long long p = r % thw;
        r /= thw;
        const int n = static_cast<int>(r % N);
        const int g = static_cast<int>(r / N);
        const uint4 v = *reinterpret_cast<const uint4*>(x + ((long long)n * thw + p) * Cp + g * 8);
        float f[8];
        unpack8(v, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = g * 8 + j;
            if (c < C) out[((long long)n * C + c) * thw + p] = f[j];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// BatchNorm finalize: per-tile partials -> batch statistics, affine scale/shift, running stats.
//   stage 1 (grid = 32-channel groups x row chunks, block = 32 channels x 32 row lanes): fp64 chunk sums
//   stage 2 (block = 32 channels x 16 chunk lanes): total, mean / var / invstd, scale / shift, running stats
// ------------------------------------------------------------------------------------------------
// One kernel: every block reduces its chunk of partial rows for 32 channels (fp64), publishes the chunk sums and takes
// a ticket; the LAST block of a channel group (all chunks published) turns them into batch statistics.  The ticket
// counters live in the caller's workspace, start at zero and are reset by the finishing block, so back-to-back calls
// on one stream can share them.  The order of the additions is fixed, so the result is deterministic.
__global__ void __launch_bounds__(1024)
bn_finalize_kernel(const float* __restrict__ part_sum, const float* __restrict__ part_sq, int part_rows, int C, int Cp,
                   int rows_per_chunk, double* __restrict__ chunk, unsigned int* __restrict__ tickets, double count,
                   const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ running_mean,
                   float* __restrict__ running_var, float momentum, float eps, float* __restrict__ scale,
                   float* __restrict__ shift, float* __restrict__ mean_out, float* __restrict__ invstd_out,
                   float4* __restrict__ bwd_table) {
    pdl_wait();
    __shared__ double sh1[32][33];
    __shared__ double sh2[32][33];
    __shared__ int last;
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    const int nchunks = gridDim.y;
    const int r0 = blockIdx.y * rows_per_chunk;
    const int r1 = min(part_rows, r0 + rows_per_chunk);
    double a = 0.0, b = 0.0;
    if (c < Cp) {
        for (int r = r0 + rl; r < r1; r += 32) {
            a += (double)part_sum[(long long)r * Cp + c];
            b += (double)part_sq[(long long)r * Cp + c];
        }
    }
    sh1[rl][cl] = a;
    sh2[rl][cl] = b;
    __syncthreads();
    // One chunk (the persistent convolution kernels write at most one partial row per SM): this block already holds
    // every row of its channels -- no chunk sums, no ticket, no second pass.  This kernel sits on the dependent chain
    // 37 times per step, its latency is all it costs.
    if (nchunks > 1) {
        if (rl == 0 && c < Cp) {
            a = b = 0.0;
            for (int r = 0; r < 32; ++r) {
                a += sh1[r][cl];
                b += sh2[r][cl];
            }
            chunk[((long long)blockIdx.y * 2 + 0) * Cp + c] = a;
            chunk[((long long)blockIdx.y * 2 + 1) * Cp + c] = b;
            __threadfence();   // publish before the ticket
        }
        __syncthreads();
        if (threadIdx.x == 0) last = (atomicAdd(&tickets[blockIdx.x], 1u) == (unsigned)(nchunks - 1));
        __syncthreads();
        if (!last) return;
        __threadfence();       // the other blocks' chunk sums are visible from here on
        a = b = 0.0;
        if (c < Cp) {
            for (int r = rl; r < nchunks; r += 32) {
                a += __ldcg(chunk + ((long long)r * 2 + 0) * Cp + c);
                b += __ldcg(chunk + ((long long)r * 2 + 1) * Cp + c);
            }
        }
        sh1[rl][cl] = a;
        sh2[rl][cl] = b;
        __syncthreads();
        if (threadIdx.x == 0) tickets[blockIdx.x] = 0u;
    }
    if (rl == 0 && c < Cp) {
        double s1 = 0.0, s2 = 0.0;
        for (int r = 0; r < 32; ++r) {
            s1 += sh1[r][cl];
            s2 += sh2[r][cl];
        }
        if (c < C) {
            const double mean = s1 / count;
            double var = s2 / count - mean * mean;
            if (var < 0.0) var = 0.0;
            const double invstd = 1.0 / sqrt(var + (double)eps);
            const float g = gamma ? gamma[c] : 1.f;
            const float bt = beta ? beta[c] : 0.f;
            scale[c] = (float)((double)g * invstd);
            shift[c] = (float)((double)bt - mean * (double)g * invstd);
            mean_out[c] = (float)mean;
            invstd_out[c] = (float)invstd;
            if (bwd_table)
                bwd_table[c] = make_float4((float)((double)g * invstd), (float)((double)bt - mean * (double)g * invstd),
                                           (float)invstd, (float)(-mean * invstd));
            if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
            if (running_var) {
                const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
                running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
            }
        } else {  // pad lanes stay exactly zero downstream
            scale[c] = 0.f;
            shift[c] = 0.f;
            mean_out[c] = 0.f;
            invstd_out[c] = 0.f;
            if (bwd_table) bwd_table[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

__global__ void bn_eval_kernel(int C, int Cp, const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ rm, const float* __restrict__ rv, float eps,
                               float* __restrict__ scale, float* __restrict__ shift) {
    pdl_wait();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= Cp) return;
    if (c < C) {
        const float invstd = rsqrtf(rv[c] + eps);
        const float g = gamma ? gamma[c] : 1.f;
        const float b = beta ? beta[c] : 0.f;
        scale[c] = g * invstd;
        shift[c] = b - rm[c] * g * invstd;
    } else {
        scale[c] = 0.f;
        shift[c] = 0.f;
    }
}

// ------------------------------------------------------------------------------------------------
// BatchNorm apply (+ second normalised branch, + residual, + ReLU)
// ------------------------------------------------------------------------------------------------
// Thread layout shared by the BN streaming kernels: a block is V = Cp/8 channel groups x R = 256/V rows.  Every thread
// keeps ONE channel group for its whole life, so the per-channel constants live in registers (no shared-memory
// lookups, no index arithmetic per element) while a block still reads/writes R whole rows = one contiguous span.
// ReLU mask of a packed bf16 pair: 0xFFFF in each half that is > 0
__device__ __forceinline__ uint32_t relu_mask2(uint32_t bf16pair) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&bf16pair);
    return __hgt2_mask(v, __float2bfloat162_rn(0.f));
}
__device__ __forceinline__ uint32_t relu2(uint32_t bf16pair) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&bf16pair);
    v = __hmax2(v, __float2bfloat162_rn(0.f));
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void load_pairs(const float* __restrict__ p, int c0, f32x2 (&out)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = f2_make(p[c0 + 2 * j], p[c0 + 2 * j + 1]);
}

// The arithmetic runs on packed fp32 pairs (one FFMA2 per two channels) and the ReLU on the packed bf16 result:
// rounding and max(., 0) commute, so the result is bit-identical to the scalar form with half the instructions.
template <bool kHasY2, bool kHasRes>
__global__ void __launch_bounds__(256)
bn_apply_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                const __nv_bfloat16* __restrict__ y2, const float* __restrict__ scale2,
                const float* __restrict__ shift2, const __nv_bfloat16* __restrict__ res,
                __nv_bfloat16* __restrict__ out, long long rows, int Cp, int R, int relu) {
    pdl_wait();
    const int V = Cp >> 3;
    const int vl = threadIdx.x % V;
    const int rl = threadIdx.x / V;
    if (rl >= R) return;
    const int c0 = vl << 3;
    f32x2 sc[4], sh[4], sc2[4], sh2[4];
    load_pairs(scale, c0, sc);
    load_pairs(shift, c0, sh);
    if (kHasY2) {
        load_pairs(scale2, c0, sc2);
        load_pairs(shift2, c0, sh2);
    }
    constexpr int U = 2;   // rows in flight per thread (4 was measured: not faster)
    const long long rstride = (long long)gridDim.x * R;
    for (long long r0 = (long long)blockIdx.x * R + rl; r0 < rows; r0 += U * rstride) {
        uint4 vy[U], vy2[U], vr[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long r = r0 + u * rstride;
            if (r < rows) {
                const long long e = r * Cp + c0;
                vy[u] = *reinterpret_cast<const uint4*>(y + e);
                if (kHasY2) vy2[u] = *reinterpret_cast<const uint4*>(y2 + e);
                if (kHasRes) vr[u] = *reinterpret_cast<const uint4*>(res + e);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long r = r0 + u * rstride;
            if (r >= rows) break;
            const uint32_t wy[4] = {vy[u].x, vy[u].y, vy[u].z, vy[u].w};
            const uint32_t wy2[4] = {vy2[u].x, vy2[u].y, vy2[u].z, vy2[u].w};
            const uint32_t wr[4] = {vr[u].x, vr[u].y, vr[u].z, vr[u].w};
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                f32x2 v = f2_fma(f2_from_bf16x2(wy[j]), sc[j], sh[j]);
                if (kHasY2) v = f2_add(v, f2_fma(f2_from_bf16x2(wy2[j]), sc2[j], sh2[j]));
                if (kHasRes) v = f2_add(v, f2_from_bf16x2(wr[j]));
                o[j] = f2_to_bf16x2(v);
                if (relu) o[j] = relu2(o[j]);
            }
            *reinterpret_cast<uint4*>(out + r * Cp + c0) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// BatchNorm backward.  dz = g * [out > 0];  s1 = sum dz, s2 = sum dz * xhat (per channel, per branch)
//   reduce : block = V channel groups x R row lanes, register accumulation, fixed-order smem reduction,
//            per-block partials [block][4][Cp]
//   final  : sums the block partials in fp64 -> sums[4][Cp] (+ dgamma/dbeta outputs)
//   apply  : dy = gamma*invstd*(dz - s1/M - xhat*s2/M)
// ------------------------------------------------------------------------------------------------
// Sums are taken against the RAW pre-activation (s2raw = sum dz*y); the final kernel turns them into
// sum dz*xhat = invstd*(s2raw - mean*s1) in fp64.  Per pair of channels: mask (FFMA2 + bf16x2 compare), two unpacks,
// one FADD2 and one FFMA2.
template <bool kHasY2>
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ out, int relu,
                     const float* __restrict__ mask_scale, const float* __restrict__ mask_shift,
                     const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ y2, long long rows, int Cp,
                     int R, float* __restrict__ partial) {
    pdl_wait();
    extern __shared__ float sm[];  // [R][4][Cp] reduction scratch
    const int V = Cp >> 3;
    const int vl = threadIdx.x % V;
    const int rl = threadIdx.x / V;
    const bool active = rl < R;
    const int c0 = vl << 3;
    f32x2 a1[4], a2[4], b2[4], msc[4], msh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) a1[j] = a2[j] = b2[j] = msc[j] = msh[j] = 0ull;
    if (relu == 2) {
        load_pairs(mask_scale, c0, msc);
        load_pairs(mask_shift, c0, msh);
    }
    if (active) {
        constexpr int U = 2;   // rows in flight per thread (4 was measured: not faster)
        const long long rstride = (long long)gridDim.x * R;
        for (long long r0 = (long long)blockIdx.x * R + rl; r0 < rows; r0 += U * rstride) {
            uint4 vg[U], vo[U], vy[U], vy2[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long r = r0 + u * rstride;
                if (r < rows) {
                    const long long e = r * Cp + c0;
                    vg[u] = *reinterpret_cast<const uint4*>(g + e);
                    if (relu == 1) vo[u] = *reinterpret_cast<const uint4*>(out + e);
                    vy[u] = *reinterpret_cast<const uint4*>(y + e);
                    if (kHasY2) vy2[u] = *reinterpret_cast<const uint4*>(y2 + e);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long r = r0 + u * rstride;
                if (r >= rows) break;
                const uint32_t wg[4] = {vg[u].x, vg[u].y, vg[u].z, vg[u].w};
                const uint32_t wo[4] = {vo[u].x, vo[u].y, vo[u].z, vo[u].w};
                const uint32_t wy[4] = {vy[u].x, vy[u].y, vy[u].z, vy[u].w};
                const uint32_t wy2[4] = {vy2[u].x, vy2[u].y, vy2[u].z, vy2[u].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const f32x2 yv = f2_from_bf16x2(wy[j]);
                    uint32_t gw = wg[j];
                    if (relu == 1) gw &= relu_mask2(wo[j]);
                    if (relu == 2) gw &= relu_mask2(f2_to_bf16x2(f2_fma(yv, msc[j], msh[j])));
                    const f32x2 gz = f2_from_bf16x2(gw);
                    a1[j] = f2_add(a1[j], gz);
                    a2[j] = f2_fma(gz, yv, a2[j]);
                    if (kHasY2) b2[j] = f2_fma(gz, f2_from_bf16x2(wy2[j]), b2[j]);
                }
            }
        }
    }
    if (active) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float lo, hi;
            f2_split(a1[j], lo, hi);
            sm[(rl * 4 + 0) * Cp + c0 + 2 * j] = lo, sm[(rl * 4 + 0) * Cp + c0 + 2 * j + 1] = hi;
            f2_split(a2[j], lo, hi);
            sm[(rl * 4 + 1) * Cp + c0 + 2 * j] = lo, sm[(rl * 4 + 1) * Cp + c0 + 2 * j + 1] = hi;
            if (kHasY2) {
                f2_split(b2[j], lo, hi);
                sm[(rl * 4 + 3) * Cp + c0 + 2 * j] = lo, sm[(rl * 4 + 3) * Cp + c0 + 2 * j + 1] = hi;
            }
        }
    }
    __syncthreads();
    const int nq = kHasY2 ? 4 : 2;
    for (int i = threadIdx.x; i < nq * Cp; i += blockDim.x) {
        const int qi = i / Cp, c = i - qi * Cp;
        float s = 0.f;
        if (qi == 2) {
            for (int r = 0; r < R; ++r) s += sm[(r * 4 + 0) * Cp + c];  // s1 of branch 2 equals s1 of branch 1
        } else {
            for (int r = 0; r < R; ++r) s += sm[(r * 4 + qi) * Cp + c];
        }
        partial[((long long)blockIdx.x * 4 + qi) * Cp + c] = s;
    }
}

// raw != 0: rows 1 / 3 of the partials are sums against the raw pre-activation and are converted here:
// sum dz*xhat = invstd * (s2raw - mean * s1).  Block = one group of 32 channels, all quantities of that group.
__global__ void __launch_bounds__(1024)
bn_bwd_final_kernel(const float* __restrict__ partial, int nblocks, int nq, int C, int Cp, int raw,
                    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ mean2,
                    const float* __restrict__ invstd2, float* __restrict__ sums, float* __restrict__ dgamma,
                    float* __restrict__ dbeta, float* __restrict__ dgamma2, float* __restrict__ dbeta2) {
    pdl_wait();
    // 8 channels (one 32-byte sector per partial row) x 128 row lanes per block: Cp/8 blocks instead of Cp/32, and a
    // thread's loads (<= 10 rows x nq quantities for the 1184-row partials of bn_bwd_reduce) are all independent, so the
    // kernel is one memory round trip plus a shared-memory tree -- it sits on the dependent chain 34 times per step.
    __shared__ double sh[4][128][9];
    __shared__ double tot[4][8];
    const int cl = threadIdx.x & 7, rl = threadIdx.x >> 3;
    const int c = blockIdx.x * 8 + cl;
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    if (c < Cp) {
        int b = rl;
        for (; b + 384 < nblocks; b += 512) {   // four rows x nq quantities in flight
#pragma unroll
            for (int qi = 0; qi < 4; ++qi) {
                if (qi < nq) {
                    const float v0 = partial[((long long)b * 4 + qi) * Cp + c];
                    const float v1 = partial[((long long)(b + 128) * 4 + qi) * Cp + c];
                    const float v2 = partial[((long long)(b + 256) * 4 + qi) * Cp + c];
                    const float v3 = partial[((long long)(b + 384) * 4 + qi) * Cp + c];
                    s[qi] += ((double)v0 + (double)v1) + ((double)v2 + (double)v3);
                }
            }
        }
        for (; b < nblocks; b += 128) {
#pragma unroll
            for (int qi = 0; qi < 4; ++qi)
                if (qi < nq) s[qi] += (double)partial[((long long)b * 4 + qi) * Cp + c];
        }
    }
#pragma unroll
    for (int qi = 0; qi < 4; ++qi) sh[qi][rl][cl] = s[qi];
    __syncthreads();
    // two-level tree in a fixed order: 256 threads sum 16 rows each, 32 threads sum the 8 partial results
    __shared__ double part[4][8][9];
    if (threadIdx.x < 256) {
        const int qi = threadIdx.x >> 6, pr = (threadIdx.x >> 3) & 7, cc = threadIdx.x & 7;
        double t = 0.0;
#pragma unroll
        for (int r = 0; r < 16; ++r) t += sh[qi][pr * 16 + r][cc];
        part[qi][pr][cc] = t;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int qi = threadIdx.x >> 3, cc = threadIdx.x & 7;
        double t = 0.0;
#pragma unroll
        for (int r = 0; r < 8; ++r) t += part[qi][r][cc];
        tot[qi][cc] = t;
    }
    __syncthreads();
    if (rl != 0 || c >= Cp) return;
    const double s1 = tot[0][cl];
    double s2 = tot[1][cl];
    if (raw) s2 = (double)invstd[c] * (s2 - (double)mean[c] * s1);
    sums[0 * Cp + c] = (float)s1;
    sums[1 * Cp + c] = (float)s2;
    if (c < C) {
        if (dbeta) dbeta[c] = (float)s1;
        if (dgamma) dgamma[c] = (float)s2;
    }
    if (nq == 4) {
        double s4 = tot[3][cl];
        if (raw) s4 = (double)invstd2[c] * (s4 - (double)mean2[c] * s1);
        sums[2 * Cp + c] = (float)s1;
        sums[3 * Cp + c] = (float)s4;
        if (c < C) {
            if (dbeta2) dbeta2[c] = (float)s1;
            if (dgamma2) dgamma2[c] = (float)s4;
        }
    }
}

// dy = k*(dz - c1 - xhat*c2) with k = gamma*invstd, c1 = mean(dz), c2 = mean(dz*xhat), xhat = (y-mu)*invstd, folded
// per channel into  dy = A*dz + B*y + D  (A = k, B = -k*c2*invstd, D = k*(c2*invstd*mu - c1)).
template <bool kHasY2>
__global__ void __launch_bounds__(256, kHasY2 ? 1 : 3)
bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ out, int relu,
                    const float* __restrict__ mask_scale, const float* __restrict__ mask_shift,
                    const __nv_bfloat16* __restrict__ y, const float* __restrict__ mean,
                    const float* __restrict__ invstd, const float* __restrict__ gamma,
                    const __nv_bfloat16* __restrict__ y2, const float* __restrict__ mean2,
                    const float* __restrict__ invstd2, const float* __restrict__ gamma2,
                    const float* __restrict__ sums, __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dy2,
                    __nv_bfloat16* __restrict__ dz, long long rows, int C, int Cp, int R, float inv_count) {
    pdl_wait();
    const int V = Cp >> 3;
    const int vl = threadIdx.x % V;
    const int rl = threadIdx.x / V;
    if (rl >= R) return;
    const int c0 = vl << 3;
    f32x2 A[4], Bc[4], D[4], A2[4], B2[4], D2[4], msc[4], msh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float a[2], b[2], d[2], a2[2] = {0.f, 0.f}, b2[2] = {0.f, 0.f}, d2[2] = {0.f, 0.f};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = c0 + 2 * j + h;
            const float gm = (c < C) ? (gamma ? gamma[c] : 1.f) : 0.f;
            const float is = invstd[c], mu = mean[c];
            const float k = gm * is;
            const float c1 = sums[c] * inv_count, c2 = sums[Cp + c] * inv_count;
            a[h] = k;
            b[h] = -k * c2 * is;
            d[h] = k * (c2 * is * mu - c1);
            if (kHasY2) {
                const float gm2 = (c < C) ? (gamma2 ? gamma2[c] : 1.f) : 0.f;
                const float is2 = invstd2[c], mu2 = mean2[c];
                const float k2 = gm2 * is2;
                const float c22 = sums[3 * Cp + c] * inv_count;
                a2[h] = k2;
                b2[h] = -k2 * c22 * is2;
                d2[h] = k2 * (c22 * is2 * mu2 - c1);
            }
        }
        A[j] = f2_make(a[0], a[1]), Bc[j] = f2_make(b[0], b[1]), D[j] = f2_make(d[0], d[1]);
        A2[j] = f2_make(a2[0], a2[1]), B2[j] = f2_make(b2[0], b2[1]), D2[j] = f2_make(d2[0], d2[1]);
        msc[j] = msh[j] = 0ull;
    }
    if (relu == 2) {
        load_pairs(mask_scale, c0, msc);
        load_pairs(mask_shift, c0, msh);
    }
    constexpr int U = 2;
    const long long rstride = (long long)gridDim.x * R;
    for (long long r0 = (long long)blockIdx.x * R + rl; r0 < rows; r0 += U * rstride) {
        uint4 vg[U], vo[U], vy[U], vy2[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long r = r0 + u * rstride;
            if (r < rows) {
                const long long e = r * Cp + c0;
                vg[u] = *reinterpret_cast<const uint4*>(g + e);
                if (relu == 1) vo[u] = *reinterpret_cast<const uint4*>(out + e);
                vy[u] = *reinterpret_cast<const uint4*>(y + e);
                if (kHasY2) vy2[u] = *reinterpret_cast<const uint4*>(y2 + e);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long r = r0 + u * rstride;
            if (r >= rows) break;
            const long long e = r * Cp + c0;
            const uint32_t wg[4] = {vg[u].x, vg[u].y, vg[u].z, vg[u].w};
            const uint32_t wo[4] = {vo[u].x, vo[u].y, vo[u].z, vo[u].w};
            const uint32_t wy[4] = {vy[u].x, vy[u].y, vy[u].z, vy[u].w};
            const uint32_t wy2[4] = {vy2[u].x, vy2[u].y, vy2[u].z, vy2[u].w};
            uint32_t o[4], o2[4], z[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const f32x2 yv = f2_from_bf16x2(wy[j]);
                uint32_t gw = wg[j];
                if (relu == 1) gw &= relu_mask2(wo[j]);
                if (relu == 2) gw &= relu_mask2(f2_to_bf16x2(f2_fma(yv, msc[j], msh[j])));   // mask from the pre-activation
                z[j] = gw;
                const f32x2 gz = f2_from_bf16x2(gw);
                o[j] = f2_to_bf16x2(f2_fma(A[j], gz, f2_fma(Bc[j], yv, D[j])));
                if (kHasY2) o2[j] = f2_to_bf16x2(f2_fma(A2[j], gz, f2_fma(B2[j], f2_from_bf16x2(wy2[j]), D2[j])));
            }
            *reinterpret_cast<uint4*>(dy + e) = make_uint4(o[0], o[1], o[2], o[3]);
            if (kHasY2) *reinterpret_cast<uint4*>(dy2 + e) = make_uint4(o2[0], o2[1], o2[2], o2[3]);
            if (dz != nullptr) *reinterpret_cast<uint4*>(dz + e) = make_uint4(z[0], z[1], z[2], z[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// MaxPool3d (kernel == stride), NDHWC, 8 channels per thread, argmax = flat window index (one byte per element: the
// window has at most 8 positions; the index tensor of C3D's pool1 is 70 MB instead of 283 MB as int32)
// ------------------------------------------------------------------------------------------------
__global__ void maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                   uint8_t* __restrict__ argmax, int N, int T, int H, int W, int Cp, int kt, int kh,
                                   int kw, int pt, int ph, int pw, int To, int Ho, int Wo) {
    pdl_wait();
    const int V = Cp >> 3;
    const long long total = (long long)N * To * Ho * Wo * V;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        const int g = static_cast<int>(r % V);
        r /= V;
        const int wo = static_cast<int>(r % Wo);
        r /= Wo;
        const int ho = static_cast<int>(r % Ho);
        r /= Ho;
        const int to = static_cast<int>(r % To);
        const int n = static_cast<int>(r / To);
        float best[8];
        int bi[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) best[j] = -INFINITY, bi[j] = -1;
        for (int a = 0; a < kt; ++a)
            for (int b = 0; b < kh; ++b)
                for (int c = 0; c < kw; ++c) {
                    const int t = to * kt + a - pt, h = ho * kh + b - ph, w = wo * kw + c - pw;
                    if (t < 0 || t >= T || h < 0 || h >= H || w < 0 || w >= W) continue;
                    float f[8];
                    unpack8(*reinterpret_cast<const uint4*>(x + ((((long long)n * T + t) * H + h) * W + w) * Cp + g * 8),
                            f);
                    const int code = (a * kh + b) * kw + c;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (f[j] > best[j] || bi[j] < 0) best[j] = f[j], bi[j] = code;
                }
        *reinterpret_cast<uint4*>(y + i * 8) = pack8(best);
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            lo |= static_cast<uint32_t>(bi[j] & 0xff) << (8 * j);
            hi |= static_cast<uint32_t>(bi[j + 4] & 0xff) << (8 * j);
        }
        *reinterpret_cast<uint2*>(argmax + i * 8) = make_uint2(lo, hi);
    }
}

// Per-thread channel sums -> one partial row per block.  Every thread of these grid-stride kernels always works on the
// same channel octet (the host makes gridDim.x * 256 a multiple of V = Cp/8), so it keeps 8 running sums; the block
// combines the threads of an octet in a fixed order: deterministic, no atomics.  partial: [gridDim.x][Cp].
__device__ __forceinline__ void block_channel_sums(const float (&acc)[8], int V, int Cp, float* __restrict__ partial) {
    __shared__ float red[256][9];
#pragma unroll
    for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = acc[j];
    __syncthreads();
    // thread t's octet is (blockIdx.x * 256 + t) % V == (t + shift) % V
    const int shift = static_cast<int>((static_cast<long long>(blockIdx.x) * 256) % V);
    for (int c = threadIdx.x; c < Cp; c += 256) {
        const int oct = c >> 3, j = c & 7;
        int t0 = oct - shift;
        if (t0 < 0) t0 += V;
        float s = 0.f;
        for (int t = t0; t < 256; t += V) s += red[t][j];
        partial[static_cast<long long>(blockIdx.x) * Cp + c] = s;
    }
}

// dz = g * [out > 0]  (C3D: ReLU after conv+bias, network.py:147-162); optional bias-gradient partial sums of dz
__global__ void __launch_bounds__(256)
relu_bwd_kernel(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ out,
                __nv_bfloat16* __restrict__ dz, long long nvec, int V, int Cp, float* __restrict__ partial) {
    pdl_wait();
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
         i += (long long)gridDim.x * blockDim.x) {
        float a[8], b[8];
        unpack8(*reinterpret_cast<const uint4*>(g + i * 8), a);
        unpack8(*reinterpret_cast<const uint4*>(out + i * 8), b);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            a[j] = b[j] > 0.f ? a[j] : 0.f;
            acc[j] += a[j];
        }
        *reinterpret_cast<uint4*>(dz + i * 8) = pack8(a);
    }
    if (partial != nullptr) block_channel_sums(acc, V, Cp, partial);
}

// per-block column sums of a bf16 [rows][Cp] tensor -> partial[block][Cp] (fp32); block = V groups x R row lanes
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const __nv_bfloat16* __restrict__ x, long long rows, int Cp, int R, float* __restrict__ partial) {
    pdl_wait();
    extern __shared__ float sm[];  // [R][Cp]
    const int V = Cp >> 3;
    const int vl = threadIdx.x % V, rl = threadIdx.x / V;
    const int c0 = vl << 3;
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 0.f;
    if (rl < R) {
        for (long long r = (long long)blockIdx.x * R + rl; r < rows; r += (long long)gridDim.x * R) {
            float f[8];
            unpack8(*reinterpret_cast<const uint4*>(x + r * Cp + c0), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] += f[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) sm[rl * Cp + c0 + j] = a[j];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < Cp; c += blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < R; ++r) s += sm[r * Cp + c];
        partial[(long long)blockIdx.x * Cp + c] = s;
    }
}

__global__ void __launch_bounds__(1024)
colsum_final_kernel(const float* __restrict__ partial, int nblocks, int C, int Cp, float* __restrict__ out) {
    pdl_wait();
    __shared__ double sh[32][33];
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    double s = 0.0;
    if (c < Cp)
        for (int b = rl; b < nblocks; b += 32) s += (double)partial[(long long)b * Cp + c];
    sh[rl][cl] = s;
    __syncthreads();
    if (rl == 0 && c < C) {
        s = 0.0;
        for (int r = 0; r < 32; ++r) s += sh[r][cl];
        out[c] = (float)s;
    }
}

__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ argmax,
                   const __nv_bfloat16* __restrict__ pooled, __nv_bfloat16* __restrict__ dx, int N, int T, int H, int W,
                   int Cp, int kt, int kh, int kw, int pt, int ph, int pw, int To, int Ho, int Wo,
                   float* __restrict__ partial) {
    pdl_wait();
    // one thread per input vector: windows do not overlap (kernel == stride), so each input element belongs to
    // exactly one window and the gradient is a gather.
    // pooled (optional) = the pooling OUTPUT when its input was a ReLU output: the selected element is that value, so
    // dz = dy * [pooled > 0] fuses the ReLU backward (network.py:147-162) without re-reading the (4-8x larger) input.
    // partial (optional): bias-gradient sums of dz -- each window's masked dy is counted once, by the thread of the
    // window's first position.
    const int V = Cp >> 3;
    const long long total = (long long)N * T * H * W * V;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        const int g = static_cast<int>(r % V);
        r /= V;
        const int w = static_cast<int>(r % W);
        r /= W;
        const int h = static_cast<int>(r % H);
        r /= H;
        const int t = static_cast<int>(r % T);
        const int n = static_cast<int>(r / T);
        const int to = (t + pt) / kt, ho = (h + ph) / kh, wo = (w + pw) / kw;
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.f;
        if (to < To && ho < Ho && wo < Wo) {
            const int code = (((t + pt) - to * kt) * kh + ((h + ph) - ho * kh)) * kw + ((w + pw) - wo * kw);
            const long long oi = ((((long long)n * To + to) * Ho + ho) * Wo + wo) * V + g;
            float f[8];
            unpack8(*reinterpret_cast<const uint4*>(dy + oi * 8), f);
            if (pooled != nullptr) {
                float m[8];
                unpack8(*reinterpret_cast<const uint4*>(pooled + oi * 8), m);
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = m[j] > 0.f ? f[j] : 0.f;
            }
            const uint2 am = *reinterpret_cast<const uint2*>(argmax + oi * 8);
            // first in-bounds position of the window (padding can cut the leading ones off)
            const int t_first = max(to * kt - pt, 0), h_first = max(ho * kh - ph, 0), w_first = max(wo * kw - pw, 0);
            const bool first = t == t_first && h == h_first && w == w_first;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int sel = static_cast<int>(((j < 4 ? am.x : am.y) >> (8 * (j & 3))) & 0xffu);
                o[j] = sel == code ? f[j] : 0.f;
                if (first) acc[j] += f[j];
            }
        }
        *reinterpret_cast<uint4*>(dx + i * 8) = pack8(o);
    }
    if (partial != nullptr) block_channel_sums(acc, V, Cp, partial);
}

int ew_blocks(long long work_items, int threads) {
    const long long want = ceil_div_ll(work_items, threads);
    return (int)std::max<long long>(1, std::min<long long>(want, (long long)sm_count() * 8));
}

constexpr int kBwdMaxBlocks = 1184;   // capacity of the partial buffers (8 blocks per SM)

// Grid-stride kernels run best with exactly as many blocks as can be resident at once: one more block than that
// starts a second wave that leaves most SMs idle (592 blocks on 3 x 148 resident slots = 1.33 waves).
template <typename F>
int resident_grid(F kernel, int threads, size_t dyn_smem, long long work_blocks, int max_blocks) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, dyn_smem) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 2;
    }
    const long long cap = std::min<long long>((long long)per_sm * sm_count(), max_blocks);
    return (int)std::max<long long>(1, std::min<long long>(cap, work_blocks));
}

}  // namespace
}  // namespace zsv

using namespace zsv;

extern "C" int zsv_repack_input(const float* x, void* out, int N, int C, int T, int H, int W, int layout,
                                int wpad_left, void* stream) {
    if (!x || !out) return fail(ZSV_ERR_BAD_ARG, "repack_input: null pointer");
    if (N < 1 || C < 1 || T < 1 || H < 1 || W < 1) return fail(ZSV_ERR_BAD_ARG, "repack_input: bad extents");
    int Cp = cpad(C), Wp = W, wl = 0;
    if (layout == ZSV_CONV_X_WFOLD) {
        if (Cp != 8) return fail(ZSV_ERR_UNSUPPORTED, "repack_input: wfold layout needs C <= 8");
        if (wpad_left < 0 || wpad_left > 8) return fail(ZSV_ERR_BAD_ARG, "repack_input: wpad_left out of range");
        Wp = W + 8;
        wl = wpad_left;
    } else if (layout != ZSV_CONV_X_NDHWC) {
        return fail(ZSV_ERR_BAD_ARG, "repack_input: unknown layout %d", layout);
    }
    const long long total = (long long)N * T * H * Wp * (Cp >> 3);
    zsv::launch(ncdhw_to_ndhwc_kernel, ew_blocks(total, 256), 256, 0, (cudaStream_t)stream, x, (__nv_bfloat16*)out, N, C, T, H,
                                                                                   W, Cp, Wp, wl);
    ZSV_LAUNCH_CHECK("ncdhw_to_ndhwc_kernel");
    return ZSV_OK;
}

extern "C" int zsv_ncdhw_to_ndhwc(const float* x, void* out, int N, int C, int T, int H, int W, void* stream) {
    return zsv_repack_input(x, out, N, C, T, H, W, ZSV_CONV_X_NDHWC, 0, stream);
}

extern "C" int zsv_ndhwc_to_ncdhw(const void* x, float* out, int N, int C, int T, int H, int W, void* stream) {
    if (!x || !out) return fail(ZSV_ERR_BAD_ARG, "ndhwc_to_ncdhw: null pointer");
    const int Cp = cpad(C);
    const long long total = (long long)N * T * H * W * (Cp >> 3);
    zsv::launch(ndhwc_to_ncdhw_kernel, ew_blocks(total, 256), 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, out, N, C,
                                                                                   T, H, W, Cp);
    ZSV_LAUNCH_CHECK("ndhwc_to_ncdhw_kernel");
    return ZSV_OK;
}

namespace {
constexpr int kFinalizeMaxChunks = 64;
}

constexpr size_t kTicketBytes = 1024;   // ticket counters at the start of the workspace (one per 32-channel group)

extern "C" size_t zsv_bn_finalize_workspace(int C) {
    return kTicketBytes + (size_t)kFinalizeMaxChunks * 2 * cpad(C) * sizeof(double);
}

extern "C" int zsv_bn_finalize(const float* part_sum, const float* part_sq, int part_rows, int C, long long count,
                               const float* gamma, const float* beta, float* running_mean, float* running_var,
                               float momentum, float eps, float* scale, float* shift, float* mean, float* invstd,
                               float* bwd_table, void* workspace, size_t workspace_bytes, void* stream) {
    if (!part_sum || !part_sq || !scale || !shift || !mean || !invstd || !workspace)
        return fail(ZSV_ERR_BAD_ARG, "bn_finalize: null pointer");
    if (part_rows < 1 || C < 1 || count < 1) return fail(ZSV_ERR_BAD_ARG, "bn_finalize: bad sizes");
    if (workspace_bytes < zsv_bn_finalize_workspace(C)) return fail(ZSV_ERR_WORKSPACE, "bn_finalize: workspace too small");
    const int Cp = cpad(C);
    cudaStream_t st = (cudaStream_t)stream;
    // up to 512 partial rows (16 loads in flight per thread) in one stage; beyond that ~128 rows per chunk spread
    // large row counts over many SMs
    int nchunks = part_rows <= 512 ? 1 : std::max(1, std::min(kFinalizeMaxChunks, ceil_div(part_rows, 128)));
    const int rows_per_chunk = ceil_div(part_rows, nchunks);
    nchunks = ceil_div(part_rows, rows_per_chunk);
    if (ceil_div(Cp, 32) * sizeof(unsigned int) > kTicketBytes) return fail(ZSV_ERR_UNSUPPORTED, "bn_finalize: too many channels");
    unsigned int* tickets = (unsigned int*)workspace;
    double* chunk = (double*)((char*)workspace + kTicketBytes);
    zsv::launch(bn_finalize_kernel, dim3(ceil_div(Cp, 32), nchunks), 1024, 0, st, part_sum, part_sq, part_rows, C, Cp, rows_per_chunk, chunk, tickets, (double)count, gamma, beta, running_mean,
        running_var, momentum, eps, scale, shift, mean, invstd, (float4*)bwd_table);
    ZSV_LAUNCH_CHECK("bn_finalize_kernel");
    return ZSV_OK;
}

extern "C" int zsv_bn_eval_scale_shift(int C, const float* gamma, const float* beta, const float* running_mean,
                                       const float* running_var, float eps, float* scale, float* shift, void* stream) {
    if (!running_mean || !running_var || !scale || !shift) return fail(ZSV_ERR_BAD_ARG, "bn_eval: null pointer");
    const int Cp = cpad(C);
    zsv::launch(bn_eval_kernel, ceil_div(Cp, 128), 128, 0, (cudaStream_t)stream, C, Cp, gamma, beta, running_mean, running_var,
                                                                        eps, scale, shift);
    ZSV_LAUNCH_CHECK("bn_eval_kernel");
    return ZSV_OK;
}

extern "C" int zsv_bn_apply(const void* y, const float* scale, const float* shift, const void* y2, const float* scale2,
                            const float* shift2, const void* residual, void* out, long long rows, int C, int relu,
                            void* stream) {
    if (!y || !scale || !shift || !out) return fail(ZSV_ERR_BAD_ARG, "bn_apply: null pointer");
    if (y2 && (!scale2 || !shift2)) return fail(ZSV_ERR_BAD_ARG, "bn_apply: second branch needs scale2/shift2");
    const int Cp = cpad(C);
    if (Cp > kMaxC) return fail(ZSV_ERR_UNSUPPORTED, "bn_apply: more than %d channels", kMaxC);
    if (rows < 1) return ZSV_OK;
    const int V = Cp >> 3;
    if (V > 256) return fail(ZSV_ERR_UNSUPPORTED, "bn_apply: channel pitch too large");
    const int R = std::max(1, 256 / V);
    const long long work = ceil_div_ll(rows, (long long)R * 2);
    cudaStream_t st = (cudaStream_t)stream;
    const __nv_bfloat16* yb = (const __nv_bfloat16*)y;
    const __nv_bfloat16* y2b = (const __nv_bfloat16*)y2;
    const __nv_bfloat16* rb = (const __nv_bfloat16*)residual;
    __nv_bfloat16* ob = (__nv_bfloat16*)out;
    if (y2 && residual)
        zsv::launch(bn_apply_kernel<true, true>, resident_grid(bn_apply_kernel<true, true>, 256, 0, work, 1 << 20), 256, 0, st, yb, scale, shift, y2b, scale2, shift2, rb, ob, rows, Cp, R, relu);
    else if (y2)
        zsv::launch(bn_apply_kernel<true, false>, resident_grid(bn_apply_kernel<true, false>, 256, 0, work, 1 << 20), 256, 0, st, yb, scale, shift, y2b, scale2, shift2, rb, ob, rows, Cp, R, relu);
    else if (residual)
        zsv::launch(bn_apply_kernel<false, true>, resident_grid(bn_apply_kernel<false, true>, 256, 0, work, 1 << 20), 256, 0, st, yb, scale, shift, y2b, scale2, shift2, rb, ob, rows, Cp, R, relu);
    else
        zsv::launch(bn_apply_kernel<false, false>, resident_grid(bn_apply_kernel<false, false>, 256, 0, work, 1 << 20), 256, 0, st, yb, scale, shift, y2b, scale2, shift2, rb, ob, rows, Cp, R, relu);
    ZSV_LAUNCH_CHECK("bn_apply_kernel");
    return ZSV_OK;
}

extern "C" size_t zsv_bn_bwd_workspace(int C) {
    const int Cp = cpad(C);
    return ((size_t)kBwdMaxBlocks * 4 * Cp + 4 * Cp) * sizeof(float);
}

extern "C" int zsv_bn_bwd(const void* g, const void* out, int relu, const float* mask_scale, const float* mask_shift,
                          const void* y, const float* mean,
                          const float* invstd, const float* gamma, const void* y2, const float* mean2,
                          const float* invstd2, const float* gamma2, void* dy, void* dy2, void* dz, float* dgamma,
                          float* dbeta, float* dgamma2, float* dbeta2, long long rows, int C, void* workspace,
                          size_t workspace_bytes, void* stream) {
    if (!g || !y || !mean || !invstd || !dy || !workspace) return fail(ZSV_ERR_BAD_ARG, "bn_bwd: null pointer");
    if (relu == 1 && !out) return fail(ZSV_ERR_BAD_ARG, "bn_bwd: relu mask needs the forward output");
    if (relu == 2 && (!mask_scale || !mask_shift)) return fail(ZSV_ERR_BAD_ARG, "bn_bwd: relu=2 needs scale/shift");
    if (relu < 0 || relu > 2) return fail(ZSV_ERR_BAD_ARG, "bn_bwd: relu must be 0, 1 or 2");
    if (relu == 2 && y2) return fail(ZSV_ERR_BAD_ARG, "bn_bwd: mask recomputation is for single-branch units");
    if (y2 && (!mean2 || !invstd2 || !dy2)) return fail(ZSV_ERR_BAD_ARG, "bn_bwd: second branch incomplete");
    const int Cp = cpad(C);
    if (Cp > kMaxC) return fail(ZSV_ERR_UNSUPPORTED, "bn_bwd: more than %d channels", kMaxC);
    if (workspace_bytes < zsv_bn_bwd_workspace(C)) return fail(ZSV_ERR_WORKSPACE, "bn_bwd: workspace too small");
    if (rows < 1) return fail(ZSV_ERR_BAD_ARG, "bn_bwd: rows < 1");
    cudaStream_t st = (cudaStream_t)stream;
    const int V = Cp >> 3;
    const int R = std::max(1, 256 / V);
    if (V > 256) return fail(ZSV_ERR_UNSUPPORTED, "bn_bwd: channel pitch too large");
    const size_t smem_r = (size_t)R * 4 * Cp * sizeof(float);
    float* partial = (float*)workspace;
    float* sums = partial + (size_t)kBwdMaxBlocks * 4 * Cp;
    const __nv_bfloat16 *gb = (const __nv_bfloat16*)g, *ob = (const __nv_bfloat16*)out, *yb = (const __nv_bfloat16*)y,
                        *y2b = (const __nv_bfloat16*)y2;
    static std::once_flag attr_once;   // entry points may be called from several threads (autograd, DataParallel)
    std::call_once(attr_once, [] {
        cudaFuncSetAttribute(bn_bwd_reduce_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cudaFuncSetAttribute(bn_bwd_reduce_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    });
    if (smem_r > 160 * 1024) return fail(ZSV_ERR_UNSUPPORTED, "bn_bwd: reduction scratch too large");
    const long long work_r = ceil_div_ll(rows, (long long)R * 4);
    const int nblocks = y2 ? resident_grid(bn_bwd_reduce_kernel<true>, 256, smem_r, work_r, kBwdMaxBlocks)
                           : resident_grid(bn_bwd_reduce_kernel<false>, 256, smem_r, work_r, kBwdMaxBlocks);
    if (y2)
        zsv::launch(bn_bwd_reduce_kernel<true>, nblocks, 256, smem_r, st, gb, ob, relu, mask_scale, mask_shift, yb, y2b, rows, Cp, R, partial);
    else
        zsv::launch(bn_bwd_reduce_kernel<false>, nblocks, 256, smem_r, st, gb, ob, relu, mask_scale, mask_shift, yb, y2b, rows, Cp, R, partial);
    ZSV_LAUNCH_CHECK("bn_bwd_reduce_kernel");
    const int nq = y2 ? 4 : 2;
    zsv::launch(bn_bwd_final_kernel, ceil_div(Cp, 8), 1024, 0, st, partial, nblocks, nq, C, Cp, 1, mean, invstd, mean2, invstd2, sums, dgamma, dbeta, dgamma2, dbeta2);
    ZSV_LAUNCH_CHECK("bn_bwd_final_kernel");
    const long long work_a = ceil_div_ll(rows, (long long)R * 2);
    const int blocks = y2 ? resident_grid(bn_bwd_apply_kernel<true>, 256, 0, work_a, 1 << 20)
                          : resident_grid(bn_bwd_apply_kernel<false>, 256, 0, work_a, 1 << 20);
    const float inv_count = (float)(1.0 / (double)rows);
    if (y2)
        zsv::launch(bn_bwd_apply_kernel<true>, blocks, 256, 0, st, gb, ob, relu, mask_scale, mask_shift, yb, mean, invstd, gamma, y2b, mean2, invstd2, gamma2, sums, (__nv_bfloat16*)dy, (__nv_bfloat16*)dy2, (__nv_bfloat16*)dz, rows, C, Cp, R, inv_count);
    else
        zsv::launch(bn_bwd_apply_kernel<false>, blocks, 256, 0, st, gb, ob, relu, mask_scale, mask_shift, yb, mean, invstd, gamma, y2b, mean2, invstd2, gamma2, sums, (__nv_bfloat16*)dy, (__nv_bfloat16*)dy2, (__nv_bfloat16*)dz, rows, C, Cp, R, inv_count);
    ZSV_LAUNCH_CHECK("bn_bwd_apply_kernel");
    return ZSV_OK;
}

extern "C" int zsv_bn_bwd_finish(const void* dz, const void* y, const float* mean, const float* invstd,
                                 const float* gamma, const float* partial, int partial_rows, void* dy, float* dgamma,
                                 float* dbeta, long long rows, int C, void* workspace, size_t workspace_bytes,
                                 void* stream) {
    if (!dz || !y || !mean || !invstd || !partial || !dy || !workspace)
        return fail(ZSV_ERR_BAD_ARG, "bn_bwd_finish: null pointer");
    if (partial_rows < 1 || rows < 1) return fail(ZSV_ERR_BAD_ARG, "bn_bwd_finish: bad sizes");
    const int Cp = cpad(C);
    if (Cp > kMaxC) return fail(ZSV_ERR_UNSUPPORTED, "bn_bwd_finish: more than %d channels", kMaxC);
    if (workspace_bytes < 4 * (size_t)Cp * sizeof(float)) return fail(ZSV_ERR_WORKSPACE, "bn_bwd_finish: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    float* sums = (float*)workspace;
    // the convolution epilogue leaves (sum dz, sum dz*y) against the RAW pre-activation; converted to sum dz*xhat in fp64
    zsv::launch(bn_bwd_final_kernel, ceil_div(Cp, 8), 1024, 0, st, partial, partial_rows, 2, C, Cp, 1, mean, invstd, nullptr, nullptr,
                                                          sums, dgamma, dbeta, nullptr, nullptr);
    ZSV_LAUNCH_CHECK("bn_bwd_final_kernel");
    const int V = Cp >> 3;
    if (V > 256) return fail(ZSV_ERR_UNSUPPORTED, "bn_bwd_finish: channel pitch too large");
    const int R = std::max(1, 256 / V);
    const int blocks = resident_grid(bn_bwd_apply_kernel<false>, 256, 0, ceil_div_ll(rows, (long long)R * 2), 1 << 20);
    const float inv_count = (float)(1.0 / (double)rows);
    zsv::launch(bn_bwd_apply_kernel<false>, blocks, 256, 0, st, (const __nv_bfloat16*)dz, nullptr, 0, nullptr, nullptr, (const __nv_bfloat16*)y, mean, invstd, gamma, nullptr,
        nullptr, nullptr, nullptr, sums, (__nv_bfloat16*)dy, nullptr, nullptr, rows, C, Cp, R, inv_count);
    ZSV_LAUNCH_CHECK("bn_bwd_apply_kernel");
    return ZSV_OK;
}

extern "C" int zsv_maxpool3d_fwd(const void* x, void* y, uint8_t* argmax, int N, int T, int H, int W, int C, int kt,
                                 int kh, int kw, int pt, int ph, int pw, void* stream) {
    if (!x || !y || !argmax) return fail(ZSV_ERR_BAD_ARG, "maxpool_fwd: null pointer");
    const int Cp = cpad(C);
    const int To = (T + 2 * pt - kt) / kt + 1, Ho = (H + 2 * ph - kh) / kh + 1, Wo = (W + 2 * pw - kw) / kw + 1;
    if (To < 1 || Ho < 1 || Wo < 1) return fail(ZSV_ERR_BAD_ARG, "maxpool_fwd: empty output");
    if (kt * kh * kw > 255) return fail(ZSV_ERR_UNSUPPORTED, "maxpool_fwd: window larger than 255 positions");
    const long long total = (long long)N * To * Ho * Wo * (Cp >> 3);
    zsv::launch(maxpool_fwd_kernel, ew_blocks(total, 256), 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, argmax, N, T, H, W, Cp, kt, kh, kw, pt, ph, pw, To, Ho, Wo);
    ZSV_LAUNCH_CHECK("maxpool_fwd_kernel");
    return ZSV_OK;
}

// grid of the backward kernels that also emit bias-gradient partial rows: total threads a multiple of V
static int bias_fused_grid(long long work_items, int V) {
    int blocks = (int)std::max<long long>(1, std::min<long long>(ceil_div_ll(work_items, 256), kBwdMaxBlocks));
    while (((long long)blocks * 256) % V != 0) --blocks;     // V <= 256 divides 256 for power-of-two pitches; else shrink
    return std::max(blocks, 1);
}

static int bias_fused_finish(const float* partial, int nblocks, int C, float* db, cudaStream_t st) {
    zsv::launch(colsum_final_kernel, ceil_div(cpad(C), 32), 1024, 0, st, partial, nblocks, C, cpad(C), db);
    ZSV_LAUNCH_CHECK("colsum_final_kernel");
    return ZSV_OK;
}

extern "C" int zsv_relu_bwd(const void* g, const void* out, void* dz, long long rows, int C, float* bias_grad,
                            void* workspace, size_t workspace_bytes, void* stream) {
    if (!g || !out || !dz) return fail(ZSV_ERR_BAD_ARG, "relu_bwd: null pointer");
    const int Cp = cpad(C), V = Cp >> 3;
    const long long nvec = rows * V;
    cudaStream_t st = (cudaStream_t)stream;
    if (bias_grad == nullptr) {
        zsv::launch(relu_bwd_kernel, ew_blocks(nvec, 1024), 256, 0, st, (const __nv_bfloat16*)g, (const __nv_bfloat16*)out,
                                                              (__nv_bfloat16*)dz, nvec, V, Cp, nullptr);
        ZSV_LAUNCH_CHECK("relu_bwd_kernel");
        return ZSV_OK;
    }
    if (!workspace || workspace_bytes < zsv_bias_grad_workspace(C)) return fail(ZSV_ERR_WORKSPACE, "relu_bwd: workspace too small");
    if ((256 % V) != 0 && V > 256) return fail(ZSV_ERR_UNSUPPORTED, "relu_bwd: channel pitch too large for the fused bias gradient");
    const int blocks = bias_fused_grid(nvec, V);
    zsv::launch(relu_bwd_kernel, blocks, 256, 0, st, (const __nv_bfloat16*)g, (const __nv_bfloat16*)out, (__nv_bfloat16*)dz, nvec, V,
                                            Cp, (float*)workspace);
    ZSV_LAUNCH_CHECK("relu_bwd_kernel");
    return bias_fused_finish((const float*)workspace, blocks, C, bias_grad, st);
}

extern "C" size_t zsv_bias_grad_workspace(int C) { return (size_t)kBwdMaxBlocks * cpad(C) * sizeof(float); }

extern "C" int zsv_bias_grad(const void* dy, float* db, long long rows, int C, void* workspace, size_t workspace_bytes,
                             void* stream) {
    if (!dy || !db || !workspace) return fail(ZSV_ERR_BAD_ARG, "bias_grad: null pointer");
    if (workspace_bytes < zsv_bias_grad_workspace(C)) return fail(ZSV_ERR_WORKSPACE, "bias_grad: workspace too small");
    const int Cp = cpad(C), V = Cp >> 3;
    if (V > 256) return fail(ZSV_ERR_UNSUPPORTED, "bias_grad: channel pitch too large");
    const int R = std::max(1, 256 / V);
    const int nblocks = (int)std::max<long long>(1, std::min<long long>(kBwdMaxBlocks, ceil_div_ll(rows, (long long)R * 4)));
    cudaStream_t st = (cudaStream_t)stream;
    zsv::launch(colsum_partial_kernel, nblocks, 256, (size_t)R * Cp * sizeof(float), st, (const __nv_bfloat16*)dy, rows, Cp, R,
                                                                                 (float*)workspace);
    ZSV_LAUNCH_CHECK("colsum_partial_kernel");
    zsv::launch(colsum_final_kernel, ceil_div(Cp, 32), 1024, 0, st, (const float*)workspace, nblocks, C, Cp, db);
    ZSV_LAUNCH_CHECK("colsum_final_kernel");
    return ZSV_OK;
}

extern "C" int zsv_maxpool3d_bwd(const void* dy, const uint8_t* argmax, const void* relu_pooled, void* dx, int N,
                                 int T, int H, int W, int C, int kt, int kh, int kw, int pt, int ph, int pw,
                                 float* bias_grad, void* workspace, size_t workspace_bytes, void* stream) {
    if (!dy || !dx || !argmax) return fail(ZSV_ERR_BAD_ARG, "maxpool_bwd: null pointer");
    const int Cp = cpad(C), V = Cp >> 3;
    const int To = (T + 2 * pt - kt) / kt + 1, Ho = (H + 2 * ph - kh) / kh + 1, Wo = (W + 2 * pw - kw) / kw + 1;
    const long long total = (long long)N * T * H * W * V;
    cudaStream_t st = (cudaStream_t)stream;
    float* partial = nullptr;
    int blocks = ew_blocks(total, 256);
    if (bias_grad != nullptr) {
        if (!workspace || workspace_bytes < zsv_bias_grad_workspace(C)) return fail(ZSV_ERR_WORKSPACE, "maxpool_bwd: workspace too small");
        if (V > 256) return fail(ZSV_ERR_UNSUPPORTED, "maxpool_bwd: channel pitch too large for the fused bias gradient");
        blocks = bias_fused_grid(total, V);
        partial = (float*)workspace;
    }
    zsv::launch(maxpool_bwd_kernel, blocks, 256, 0, st, (const __nv_bfloat16*)dy, argmax, (const __nv_bfloat16*)relu_pooled,
                                               (__nv_bfloat16*)dx, N, T, H, W, Cp, kt, kh, kw, pt, ph, pw, To, Ho, Wo, partial);
    ZSV_LAUNCH_CHECK("maxpool_bwd_kernel");
    if (bias_grad != nullptr) return bias_fused_finish(partial, blocks, C, bias_grad, st);
    return ZSV_OK;
}

// ------------------------------------------------------------------------------------------------
// Input pipeline stage (auxiliary/transforms.py:41-56 get_transform): uint8 THWC frames ->
//   ToFloatTensorInZeroOne  (x/255 - 1)/2                                   transforms.py:116-117
//   Resize(128)             bilinear, align_corners=False, scale = 128/min(H,W)   transforms.py:99-108
//   Center/RandomCrop(112)  window origin (i, j) per clip                   transforms.py:132-165
//   RandomHorizontalFlip    per clip                                        transforms.py:189-195
// fused with the layout conversion of zsv_repack_input: one pass from the decoded frames to the bf16 W-folded tensor
// the first convolution reads.  One thread per output position (3 channels + 5 zero lanes = one 16-byte store).
// ------------------------------------------------------------------------------------------------
namespace zsv {
namespace {
__global__ void clip_transform_kernel(const uint8_t* __restrict__ frames, __nv_bfloat16* __restrict__ out, int N, int T,
                                      int Hs, int Ws, int Hr, int Wr, float rscale, int crop, int Wp, int wl,
                                      const int32_t* __restrict__ crop_ij, const uint8_t* __restrict__ flip) {
    pdl_wait();
    const long long total = (long long)N * T * crop * Wp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        const int wp = static_cast<int>(r % Wp);
        r /= Wp;
        const int h = static_cast<int>(r % crop);
        r /= crop;
        const int t = static_cast<int>(r % T);
        const int n = static_cast<int>(r / T);
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const int w = wp - wl;
        if (w >= 0 && w < crop) {
            const int wsrc = flip != nullptr && flip[n] ? crop - 1 - w : w;
            const int yo = crop_ij[2 * n] + h, xo = crop_ij[2 * n + 1] + wsrc;   // position in the resized frame
            // PyTorch upsample_bilinear2d, align_corners = false, scale given: src = (dst + 0.5) / scale - 0.5, clamped at 0
            float sy = rscale * (yo + 0.5f) - 0.5f, sx = rscale * (xo + 0.5f) - 0.5f;
            sy = sy < 0.f ? 0.f : sy;
            sx = sx < 0.f ? 0.f : sx;
            const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
            const int y1 = y0 + (y0 < Hs - 1 ? 1 : 0), x1 = x0 + (x0 < Ws - 1 ? 1 : 0);
            const float ly = sy - y0, lx = sx - x0;
            const float hy = 1.f - ly, hx = 1.f - lx;
            const uint8_t* f = frames + ((long long)n * T + t) * Hs * Ws * 3;
            (void)Hr;
            (void)Wr;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float p00 = (f[((long long)y0 * Ws + x0) * 3 + c] / 255.f - 1.0f) / 2.0f;
                const float p01 = (f[((long long)y0 * Ws + x1) * 3 + c] / 255.f - 1.0f) / 2.0f;
                const float p10 = (f[((long long)y1 * Ws + x0) * 3 + c] / 255.f - 1.0f) / 2.0f;
                const float p11 = (f[((long long)y1 * Ws + x1) * 3 + c] / 255.f - 1.0f) / 2.0f;
                v[c] = hy * (hx * p00 + lx * p01) + ly * (hx * p10 + lx * p11);
            }
        }
        *reinterpret_cast<uint4*>(out + i * 8) = pack8(v);
    }
}
}  // namespace
}  // namespace zsv

extern "C" int zsv_clip_transform(const uint8_t* frames, void* out, int N, int T, int Hs, int Ws, int resize_short,
                                  int crop, const int32_t* crop_ij, const uint8_t* flip, int wpad_left, void* stream) {
    if (!frames || !out || !crop_ij) return fail(ZSV_ERR_BAD_ARG, "clip_transform: null pointer");
    if (N < 1 || T < 1 || Hs < 1 || Ws < 1 || resize_short < 1 || crop < 1) return fail(ZSV_ERR_BAD_ARG, "clip_transform: bad extents");
    if (wpad_left < 0 || wpad_left > 8) return fail(ZSV_ERR_BAD_ARG, "clip_transform: wpad_left out of range");
    // transforms.py:103-105: scale = float(size) / min(h, w); F.interpolate(scale_factor=scale) -> floor(in * scale)
    const double scale = (double)resize_short / (double)std::min(Hs, Ws);
    const int Hr = (int)floor((double)Hs * scale), Wr = (int)floor((double)Ws * scale);
    if (Hr < crop || Wr < crop) return fail(ZSV_ERR_BAD_ARG, "clip_transform: crop %d larger than the resized frame %dx%d", crop, Hr, Wr);
    const float rscale = (float)(1.0 / scale);
    const int Wp = crop + 8;
    const long long total = (long long)N * T * crop * Wp;
    zsv::launch(zsv::clip_transform_kernel, ew_blocks(total, 256), 256, 0, (cudaStream_t)stream, frames, (__nv_bfloat16*)out, N, T, Hs, Ws, Hr, Wr, rscale, crop, Wp, wpad_left, crop_ij, flip);
    ZSV_LAUNCH_CHECK("clip_transform_kernel");
    return ZSV_OK;
}
