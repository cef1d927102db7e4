// Embedding head and zero-shot nearest-class search.
//
// Head (network.py:595-596, MLP network.py:603-618): mean over (T,H,W) -> Linear -> ReLU -> Linear ->
// F.normalize, forward and analytic backward, all fp32 on CUDA cores: with B ~ 22 rows these are
// launch/latency-bound weight-streaming kernels, not tensor-core work.
// Loss (main.py:130,179): MSELoss(mean).
// Nearest class (main.py:183, main.py:321-322): scipy cdist(...,'cosine') + argmin / argsort[:, :k],
// restated in fp64 with scipy's exact operation order so that indices are bit-identical.
#include <algorithm>
#include <stdlib.h>
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
    pdl_wait();
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
    pdl_wait();
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
    pdl_wait();
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
    pdl_wait();
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
    pdl_wait();
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
// values are exact in fp64, so one DFMA per (row, class, k) equals scipy's mul + add bit for bit.
// cosine = dot / (nu*nv), clipped to +-1, d = 1 - cosine.
//
// The dot products are a register-tiled fp64 contraction (the FP64 pipe is the bound: N*C*D DFMAs at 64 per clock
// and SM).  A warp owns R embedding rows x 128 class slots (lane l: classes 2l, 2l+1, 64+2l, 64+2l+1) for ONE of scipy's two
// accumulation lanes: the even-k warp and the odd-k warp of a tile run side by side and meet at the end
// (total = even + odd), so a thread carries R*4 accumulators and the operand loads of the next k overlap the DFMAs
// of the current one.  A block is 2 (parity) x G (row groups) x TC (class tiles) warps.  Both operands are staged
// k-major as doubles in shared memory, 16 k at a time: per k a thread reads its four classes with two 16-byte loads
// and its R rows with R/2 broadcast loads, for 4R DFMAs; the next chunk travels global -> registers meanwhile.
// Norms ride along: thread i squares class slot i (and row i) of every staged chunk, in scipy's order.
// The staging buffers alternate, so there is one block barrier per chunk.
//   evaluation (main.py:321, thousands of rows): a block of G = 2 or 3 row groups walks all class tiles itself.
//   train-time batch (main.py:183, ~22 rows x several hundred classes): the class tiles of one 24-row block are spread
//   over a cluster of up to 8 CTAs; each writes its keys into the leader's table through distributed shared memory.
// Distances are staged in shared memory [rows][C] as order-preserving 64-bit keys (NaN last); one warp per row then
// extracts the k smallest with three warp-wide integer min reductions per pick (high word, low word, index: lowest
// index wins ties).
// ------------------------------------------------------------------------------------------------
constexpr int kNK = 16;    // k-chunk staged in shared memory
constexpr int kCT = 128;   // class slots of one warp tile (32 lanes x 4)

// four consecutive fp32 of row `row` starting at column k (zeros past the end of the row / table)
__device__ __forceinline__ float4 ld_row4(const float* __restrict__ base, long long row, int nrows, int k, int D,
                                          bool vec) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < nrows && k < D) {
        const float* p = base + row * D + k;
        if (vec) {
            v = __ldg(reinterpret_cast<const float4*>(p));      // D % 4 == 0 and a 16-byte aligned base
        } else {
            v.x = p[0];
            if (k + 1 < D) v.y = p[1];
            if (k + 2 < D) v.z = p[2];
            if (k + 3 < D) v.w = p[3];
        }
    }
    return v;
}

// order-preserving map double -> uint64 (every NaN becomes the largest key but one; all-ones marks "taken")
__device__ __forceinline__ unsigned long long dist_key(double d) {
    if (d != d) return 0xFFFFFFFFFFFFFFFEull;
    const unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_dist(unsigned long long key) {
    if (key == 0xFFFFFFFFFFFFFFFEull) return __longlong_as_double(0x7FF8000000000000ll);
    const unsigned long long b = (key >> 63) ? (key & 0x7FFFFFFFFFFFFFFFull) : ~key;
    return __longlong_as_double((long long)b);
}

// class slot (within a 128-slot warp tile) of accumulator column c of a lane: {2l, 2l+1, 64+2l, 64+2l+1}, so that each
// of the two 16-byte operand loads of a k step is contiguous across the warp (no bank conflicts)
__device__ __forceinline__ int nearest_slot(int lane, int c) { return 2 * lane + (c & 1) + 64 * (c >> 1); }

template <int R>
__device__ __forceinline__ void nearest_kstep(const double* __restrict__ qc, const double* __restrict__ qe,
                                              double (&acc)[R][4]) {
    double c[4], e[R];
    *reinterpret_cast<double2*>(&c[0]) = *reinterpret_cast<const double2*>(qc);
    *reinterpret_cast<double2*>(&c[2]) = *reinterpret_cast<const double2*>(qc + 64);
#pragma unroll
    for (int r = 0; r < R; r += 2) *reinterpret_cast<double2*>(&e[r]) = *reinterpret_cast<const double2*>(qe + r);
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[r][j] = fma(e[r], c[j], acc[r][j]);
}

template <int R, int G, int TC, bool CL>
__global__ void __launch_bounds__(64 * G * TC, CL ? 1 : 384 / (64 * G * TC))
nearest_kernel(const float* __restrict__ emb, const float* __restrict__ cls, int N, int C, int D, int k,
               int64_t* __restrict__ idx_out, double* __restrict__ dist_out, int vec, int ncta) {
    pdl_wait();
    constexpr int NT = 64 * G * TC;      // threads
    constexpr int RB = G * R;            // rows per block
    constexpr int CS = TC * kCT;         // class slots per pass
    constexpr int CP = CS + 2;           // pitch of a staged k-row of classes (keeps 16-byte alignment, spreads banks)
    constexpr int NCF = (CS * 4 + NT - 1) / NT;   // float4 prefetch registers for the class chunk
    constexpr int NCN = (CS + NT - 1) / NT;       // class slots whose norm a thread accumulates
    constexpr int SB = kNK * (CP + RB);           // doubles of one staging buffer
    extern __shared__ double smd[];
    double* dist = smd;                              // [RB][C] distance keys (cluster mode: only the leader's is used)
    double* nrm_e = dist + (size_t)RB * C;           // [RB]
    double* nrm_c = nrm_e + RB;                      // [CS]
    double* stage = nrm_c + CS;                      // 2 x { [kNK][CP] classes, [kNK][RB] rows }
    double* part = stage + 2 * SB;                   // cluster mode: [RB][CS] odd-lane partial sums (else they sit in dist)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int par = warp & 1;                        // scipy's accumulation lane of this warp: even or odd k
    const int g = (warp >> 1) % G, t = (warp >> 1) / G;
    const int rank = CL ? (int)cluster_ctarank() : 0;
    const long long i0 = (long long)(CL ? blockIdx.x / ncta : blockIdx.x) * RB;
    const int Deven = D & ~1;
    const int nchunks = (D + kNK - 1) / kNK;
    const int ntiles = (C + CS - 1) / CS;
    const int npass = CL ? (ntiles + ncta - 1) / ncta : ntiles;
    // staging: float4 number n of this thread covers class slot (tid + NT n) / 4, k offset 4 * (tid & 3)
    const int fq = tid & 3;
    const bool erow = tid < RB * 4;                  // this thread also stages a row float4
    const int er = tid >> 2;
    const uint32_t dist_leader = CL ? mapa_shared(smem_u32(dist), 0) : 0u;

    double ee0 = 0.0, ee1 = 0.0;                     // squared norm of row `tid` (threads < RB), scipy's two lanes
    for (int pass = 0; pass < npass; ++pass) {
        const int j0 = (CL ? pass * ncta + rank : pass) * CS;
        if (j0 >= C) break;                          // (cluster mode: this CTA has no tile in the last round)
        double acc[R][4];
        double cc0[NCN], cc1[NCN];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
#pragma unroll
        for (int n = 0; n < NCN; ++n) cc0[n] = cc1[n] = 0.0;

        float4 pc[NCF], pe;
        auto fetch = [&](int k0) {
#pragma unroll
            for (int n = 0; n < NCF; ++n) {
                const int i = tid + NT * n;
                pc[n] = i < CS * 4 ? ld_row4(cls, j0 + (i >> 2), C, k0 + 4 * fq, D, vec) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            pe = erow ? ld_row4(emb, i0 + er, N, k0 + 4 * fq, D, vec) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        auto put = [&](double* buf) {
#pragma unroll
            for (int n = 0; n < NCF; ++n) {
                const int i = tid + NT * n;
                if (i < CS * 4) {
                    double* d = buf + (4 * fq) * CP + (i >> 2);
                    d[0] = (double)pc[n].x, d[CP] = (double)pc[n].y, d[2 * CP] = (double)pc[n].z, d[3 * CP] = (double)pc[n].w;
                }
            }
            if (erow) {
                double* d = buf + kNK * CP + (4 * fq) * RB + er;
                d[0] = (double)pe.x, d[RB] = (double)pe.y, d[2 * RB] = (double)pe.z, d[3 * RB] = (double)pe.w;
            }
        };
        fetch(0);
        put(stage);
        if (nchunks > 1) fetch(kNK);
        __syncthreads();
        for (int ch = 0; ch < nchunks; ++ch) {
            const int k0 = ch * kNK;
            // chunk ch + 1 (in registers) goes to the other buffer, chunk ch + 2 starts its way from global memory,
            // then chunk ch is contracted: one barrier per chunk
            if (ch + 1 < nchunks) put(stage + ((ch + 1) & 1) * SB);
            if (ch + 2 < nchunks) fetch(k0 + 2 * kNK);
            const double* s_c = stage + (ch & 1) * SB;
            const double* s_e = s_c + kNK * CP;
            const int kmax = min(kNK, Deven - k0);   // kNK is even: chunk-local parity == global parity
            const double* qc = s_c + par * CP + t * kCT + 2 * lane;
            const double* qe = s_e + par * RB + g * R;
            if (kmax == kNK) {
#pragma unroll
                for (int i = 0; i < kNK / 2; ++i) nearest_kstep<R>(qc + 2 * i * CP, qe + 2 * i * RB, acc);
            } else {
                for (int kk = par; kk < kmax; kk += 2) nearest_kstep<R>(qc + (kk - par) * CP, qe + (kk - par) * RB, acc);
            }
            // norms of class slot tid (+ NT n) and, on the first pass, of row tid
#pragma unroll
            for (int n = 0; n < NCN; ++n) {
                const int slot = tid + NT * n;
                if (slot < CS) {
                    const double* q = s_c + slot;
                    for (int kk = 0; kk < kmax; kk += 2) {
                        const double v0 = q[kk * CP], v1 = q[(kk + 1) * CP];
                        cc0[n] = fma(v0, v0, cc0[n]);
                        cc1[n] = fma(v1, v1, cc1[n]);
                    }
                }
            }
            if (pass == 0 && tid < RB) {
                for (int kk = 0; kk < kmax; kk += 2) {
                    const double v0 = s_e[kk * RB + tid], v1 = s_e[(kk + 1) * RB + tid];
                    ee0 = fma(v0, v0, ee0);
                    ee1 = fma(v1, v1, ee1);
                }
            }
            __syncthreads();
        }
        // the last chunk is still staged: it holds the odd tail column (D odd) at row Deven - k0
        const double* s_c = stage + ((nchunks - 1) & 1) * SB;
        const double* s_e = s_c + kNK * CP;
        const int kt = Deven - (nchunks - 1) * kNK;
        const bool tail = Deven < D;
#pragma unroll
        for (int n = 0; n < NCN; ++n) {
            const int slot = tid + NT * n;
            if (slot < CS) {
                double s = cc0[n] + cc1[n];
                if (tail) {
                    const double v = s_c[kt * CP + slot];
                    s = fma(v, v, s);
                }
                nrm_c[slot] = sqrt(s);
            }
        }
        if (pass == 0 && tid < RB) {
            double s = ee0 + ee1;
            if (tail) {
                const double v = s_e[kt * RB + tid];
                s = fma(v, v, s);
            }
            nrm_e[tid] = sqrt(s);
        }
        // the odd-k warp hands its sums to the even-k warp of the same tile
        double* pbase = CL ? part + (size_t)(g * R) * CS + t * kCT : dist + (size_t)(g * R) * C + j0 + t * kCT;
        const int ppitch = CL ? CS : C;
        if (par) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int sl = nearest_slot(lane, c);
                if (j0 + t * kCT + sl < C) {
#pragma unroll
                    for (int r = 0; r < R; ++r) pbase[(size_t)r * ppitch + sl] = acc[r][c];
                }
            }
        }
        double ct[4], et[R];
        if (tail && !par) {
#pragma unroll
            for (int c = 0; c < 4; ++c) ct[c] = s_c[kt * CP + t * kCT + nearest_slot(lane, c)];
#pragma unroll
            for (int r = 0; r < R; ++r) et[r] = s_e[kt * RB + g * R + r];
        }
        __syncthreads();
        if (!par) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int sl = nearest_slot(lane, c);
                const int slot = t * kCT + sl;
                const int j = j0 + slot;
                if (j < C) {
                    const double nv = nrm_c[slot];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        double dot = acc[r][c] + pbase[(size_t)r * ppitch + sl];          // even + odd
                        if (tail) dot = fma(et[r], ct[c], dot);
                        double cosine = dot / (nrm_e[g * R + r] * nv);
                        if (fabs(cosine) > 1.0) cosine = copysign(1.0, cosine);
                        const unsigned long long key = dist_key(1.0 - cosine);
                        if (CL) {
                            const uint32_t a = dist_leader + (uint32_t)(((size_t)(g * R + r) * C + j) * 8);
                            asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(a), "l"(key) : "memory");
                        } else {
                            *reinterpret_cast<unsigned long long*>(dist + (size_t)(g * R + r) * C + j) = key;
                        }
                    }
                }
            }
        }
    }
    if (CL) {
        cluster_sync_all();            // every CTA's keys have landed in the leader's table
        if (rank != 0) return;
    } else {
        __syncthreads();
    }
    // selection: warp w handles rows w, w + #warps, ...
    constexpr unsigned kFull = 0xffffffffu;
    constexpr unsigned long long kTaken = 0xFFFFFFFFFFFFFFFFull;
    for (int r = warp; r < RB; r += NT / 32) {
        if (i0 + r >= N) continue;
        unsigned long long* dr = reinterpret_cast<unsigned long long*>(dist) + (size_t)r * C;
        if (C <= 128) {
            // the lane's (at most four) keys, classes lane, lane + 32, ..., sorted once in registers by (key, index);
            // every pick then compares the lanes' heads
            unsigned long long v[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) v[m] = lane + 32 * m < C ? dr[lane + 32 * m] : kTaken;
            int id[4] = {lane, lane + 32, lane + 64, lane + 96};
            auto cx = [&](int a, int b) {
                if (v[b] < v[a] || (v[b] == v[a] && id[b] < id[a])) {
                    const unsigned long long tv = v[a];
                    v[a] = v[b], v[b] = tv;
                    const int ti = id[a];
                    id[a] = id[b], id[b] = ti;
                }
            };
            cx(0, 1), cx(2, 3), cx(0, 2), cx(1, 3), cx(1, 2);
            for (int sel = 0; sel < k; ++sel) {
                const unsigned hi = (unsigned)(v[0] >> 32), lo = (unsigned)v[0];
                const unsigned mhi = __reduce_min_sync(kFull, hi);
                const unsigned mlo = __reduce_min_sync(kFull, hi == mhi ? lo : 0xFFFFFFFFu);
                const bool mine = hi == mhi && lo == mlo;
                const int widx = (int)__reduce_min_sync(kFull, mine ? (unsigned)id[0] : 0x7FFFFFFFu);
                if (mine && id[0] == widx) {
                    idx_out[(i0 + r) * k + sel] = widx;
                    if (dist_out) dist_out[(i0 + r) * k + sel] = key_dist(v[0]);
                    v[0] = v[1], v[1] = v[2], v[2] = v[3], v[3] = kTaken;
                    id[0] = id[1], id[1] = id[2], id[2] = id[3];
                }
            }
            continue;
        }
        for (int sel = 0; sel < k; ++sel) {
            unsigned long long best = kTaken;
            int bidx = 0x7FFFFFFF;
            for (int j = lane; j < C; j += 32) {
                const unsigned long long v = dr[j];
                if (v < best) best = v, bidx = j;      // ascending j: the first of equal keys stays
            }
            const unsigned hi = (unsigned)(best >> 32), lo = (unsigned)best;
            const unsigned mhi = __reduce_min_sync(kFull, hi);
            const unsigned mlo = __reduce_min_sync(kFull, hi == mhi ? lo : 0xFFFFFFFFu);
            const bool mine = hi == mhi && lo == mlo;
            const int widx = (int)__reduce_min_sync(kFull, mine ? (unsigned)bidx : 0x7FFFFFFFu);
            if (mine && bidx == widx) {
                idx_out[(i0 + r) * k + sel] = widx;
                if (dist_out) dist_out[(i0 + r) * k + sel] = key_dist(best);
                dr[widx] = kTaken;
            }
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
    zsv::launch(pool_kernel, dim3(ceil_div(C, 128), B), 128, 0, st, (const __nv_bfloat16*)feat, pooled, B, P, C, cpad(C));
    ZSV_LAUNCH_CHECK("pool_kernel");
    int rc = linear_forward(pooled, w1, b1, hidden, B, C, Hd, 1, nullptr, 0, st);
    if (rc) return rc;
    // raw projection goes to emb, then normalised in place
    rc = linear_forward(hidden, w2, b2, emb, B, Hd, E, 0, nullptr, 0, st);
    if (rc) return rc;
    zsv::launch(normalize_fwd_kernel, ceil_div(B * 32, 128), 128, 0, st, emb, emb, onorm, B, E, eps);
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
    zsv::launch(normalize_bwd_kernel, ceil_div(B * 32, 128), 128, 0, st, demb, emb, onorm, dout, B, E, eps);
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
        zsv::launch(pool_bwd_kernel, (int)std::min<long long>(ceil_div_ll(total, 256), 148 * 8), 256, 0, st, dpooled, (__nv_bfloat16*)dfeat, B, P, C, cpad(C));
        ZSV_LAUNCH_CHECK("pool_bwd_kernel");
    }
    return ZSV_OK;
}

extern "C" int zsv_l2norm_fwd(const float* o, float* emb, float* onorm, int B, int E, float eps, void* stream) {
    if (!o || !emb || !onorm) return fail(ZSV_ERR_BAD_ARG, "l2norm_fwd: null pointer");
    zsv::launch(normalize_fwd_kernel, ceil_div(B * 32, 128), 128, 0, (cudaStream_t)stream, o, emb, onorm, B, E, eps);
    ZSV_LAUNCH_CHECK("normalize_fwd_kernel");
    return ZSV_OK;
}

extern "C" int zsv_l2norm_bwd(const float* demb, const float* emb, const float* onorm, float* dout, int B, int E,
                              float eps, void* stream) {
    if (!demb || !emb || !onorm || !dout) return fail(ZSV_ERR_BAD_ARG, "l2norm_bwd: null pointer");
    zsv::launch(normalize_bwd_kernel, ceil_div(B * 32, 128), 128, 0, (cudaStream_t)stream, demb, emb, onorm, dout, B, E, eps);
    ZSV_LAUNCH_CHECK("normalize_bwd_kernel");
    return ZSV_OK;
}

extern "C" int zsv_mse_fwd_bwd(const float* emb, const float* target, int B, int E, float grad_scale, float* loss,
                               float* demb, void* stream) {
    if (!emb || !target) return fail(ZSV_ERR_BAD_ARG, "mse: null pointer");
    zsv::launch(mse_kernel, 1, 1024, 0, (cudaStream_t)stream, emb, target, B * E, grad_scale, loss, demb);
    ZSV_LAUNCH_CHECK("mse_kernel");
    return ZSV_OK;
}

extern "C" int zsv_nearest_class(const float* emb, const float* cls, int N, int C, int D, int k, int64_t* idx_out,
                                 double* dist_out, void* stream) {
    if (!emb || !cls || !idx_out) return fail(ZSV_ERR_BAD_ARG, "nearest_class: null pointer");
    if (N < 0 || C < 1 || D < 1 || k < 1 || k > 8 || k > C) return fail(ZSV_ERR_BAD_ARG, "nearest_class: bad sizes");
    if (N == 0) return ZSV_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int vec = (D % 4 == 0) && (reinterpret_cast<uintptr_t>(emb) % 16 == 0) && (reinterpret_cast<uintptr_t>(cls) % 16 == 0);
    auto smem_of = [&](int RB, bool cl) {
        return sizeof(double) * ((size_t)RB * C + RB + kCT + 2 * (size_t)kNK * (kCT + 2 + RB) + (cl ? (size_t)RB * kCT : 0));
    };
    // many rows (evaluation, main.py:321): a block owns 8 G rows and walks the class tiles, G = 2..3 whichever divides
    // the blocks over the SMs with the least idle tail.  Few rows (the train-time batch, main.py:183) or a class table
    // whose distance rows would not fit beside the staging buffers: the class tiles of a row block are spread over a
    // cluster of up to 8 CTAs that write their keys into the leader's table (distributed shared memory).
    int G = 0;
    if (const char* e = getenv("ZSV_DEBUG_NEAREST_G")) {
        G = atoi(e);
        if (G < 2 || G > 3 || smem_of(8 * G, false) > 100 * 1024) G = 0;
    } else if (N >= 1024) {
        double best = 0.0;
        for (int g = 3; g >= 2; --g) {
            if (smem_of(8 * g, false) > 100 * 1024) continue;
            const double per_sm = (double)ceil_div(N, 8 * g) / sm_count();
            const double eff = per_sm / ceil(per_sm) * (g == 3 ? 1.0 : 0.97);   // fewer rows per staged class tile cost a little
            if (eff > best) best = eff, G = g;
        }
    }
    static std::once_flag attr_once;
    std::call_once(attr_once, [] {
        cudaFuncSetAttribute(nearest_kernel<8, 3, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(nearest_kernel<8, 2, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(nearest_kernel<8, 3, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(nearest_kernel<8, 1, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    });
    if (G) {
        const size_t smem = smem_of(8 * G, false);
        if (G == 3)
            zsv::launch(nearest_kernel<8, 3, 1, false>, ceil_div(N, 24), 192, smem, st, emb, cls, N, C, D, k, idx_out, dist_out, vec, 1);
        else
            zsv::launch(nearest_kernel<8, 2, 1, false>, ceil_div(N, 16), 128, smem, st, emb, cls, N, C, D, k, idx_out, dist_out, vec, 1);
    } else {
        const int g = smem_of(24, true) <= 200 * 1024 ? 3 : 1;
        const size_t smem = smem_of(8 * g, true);
        if (smem > 200 * 1024)
            return fail(ZSV_ERR_UNSUPPORTED, "nearest_class: class table too large for shared memory (%zu bytes)", smem);
        const int ncta = std::min(8, ceil_div(C, kCT));
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(ncta * ceil_div(N, 8 * g)), cfg.blockDim = dim3(64 * g), cfg.dynamicSmemBytes = smem, cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = ncta, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
        pdl_attribute(&attr[1]);
        cfg.attrs = attr, cfg.numAttrs = 2;
        cudaError_t e = g == 3 ? cudaLaunchKernelEx(&cfg, nearest_kernel<8, 3, 1, true>, emb, cls, N, C, D, k, idx_out, dist_out, vec, ncta)
                               : cudaLaunchKernelEx(&cfg, nearest_kernel<8, 1, 1, true>, emb, cls, N, C, D, k, idx_out, dist_out, vec, ncta);
        if (e != cudaSuccess) return fail(ZSV_ERR_CUDA, "launch of nearest_kernel (cluster) failed: %s", cudaGetErrorString(e));
    }
    ZSV_LAUNCH_CHECK("nearest_kernel");
    return ZSV_OK;
}
