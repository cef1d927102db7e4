// Optimizer step of the training iteration (main.py:131,200-203: torch.optim.Adam) on sm_100a.
//
// Two kernels, both HBM-bound (16 B read + 12 B written per parameter, fp32 state):
//   * adam_plain_kernel      -- multi-tensor Adam over arbitrary fp32 tensors (BatchNorm gamma / beta, the MLP head, the
//                               W-folded first convolution), one 2048-element unit per block iteration;
//   * adam_pack_tiled_kernel -- Adam over convolution weights FUSED with the bf16 re-pack the next forward needs: a block
//                               owns 16 output channels x one chunk of input channels x all taps (16 contiguous runs of the
//                               fp32 master weight), updates p / m / v in place, stages the UPDATED weights as bf16 in shared
//                               memory and writes the fprop image [tap][Cout][kpitch] and the dgrad image [tap][Cin][copitch]
//                               with 16-byte stores.  The separate re-pack pass (re-reading 127 MB of fp32 weights per step
//                               for R(2+1)D-18) disappears.
// The step counter and (optionally) the learning rate are DEVICE scalars read at run time, so the update can be captured
// into a CUDA graph and an LR schedule (main.py:133,374: MultiStepLR) still reaches the replays.
#include <algorithm>
#include <string.h>

#include <cuda_bf16.h>

#include "zsv_internal.h"
#include "zsv_ptx.cuh"

namespace zsv {
namespace {

struct HyperDev {
    float lr, beta1, beta2, eps, weight_decay, grad_scale;
    const float* lr_dev;
};

struct AdamCoef {
    float step_size, bc2_sqrt, b1c, b2, b2c, eps, wd, gs;
};

__device__ __forceinline__ AdamCoef adam_coef(const HyperDev& h, const float* step) {
    const float t = *step;
    const float lr = h.lr_dev != nullptr ? *h.lr_dev : h.lr;
    AdamCoef c;
    c.step_size = lr / (1.f - powf(h.beta1, t));
    c.bc2_sqrt = sqrtf(1.f - powf(h.beta2, t));
    c.b1c = 1.f - h.beta1, c.b2 = h.beta2, c.b2c = 1.f - h.beta2, c.eps = h.eps, c.wd = h.weight_decay, c.gs = h.grad_scale;
    return c;
}

//   g' = g*grad_scale + wd*p ; m += (1-b1)*(g'-m) ; v = b2*v + (1-b2)*g'^2 ; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
__device__ __forceinline__ void adam_update(const AdamCoef& c, float& p, float g, float& m, float& v) {
    g *= c.gs;
    if (c.wd != 0.f) g = fmaf(c.wd, p, g);
    m = m + c.b1c * (g - m);                    // lerp, like torch
    v = c.b2 * v + c.b2c * g * g;
    const float denom = sqrtf(v) / c.bc2_sqrt + c.eps;
    p -= c.step_size * (m / denom);
}

// ---- plain multi-tensor Adam ---------------------------------------------------------------------------------------
constexpr int kPlainItems = 56;
constexpr int kPlainUnit = 2048;       // elements per block iteration: 256 threads x 8
struct PlainItem {
    float* p;
    const float* g;
    float* m;
    float* v;
    const float* step;
    long long n;
    int32_t unit_start;
    int32_t pad_;
};
struct PlainBatch {
    int32_t n, units;
    HyperDev h;
    PlainItem it[kPlainItems];
};

__global__ void __launch_bounds__(256)
adam_plain_kernel(const __grid_constant__ PlainBatch B) {
    pdl_wait();
    for (int unit = blockIdx.x; unit < B.units; unit += gridDim.x) {
        int lo = 0, hi = B.n - 1;   // last item with unit_start <= unit
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (B.it[mid].unit_start <= unit) lo = mid;
            else hi = mid - 1;
        }
        const PlainItem& I = B.it[lo];
        const AdamCoef c = adam_coef(B.h, I.step);
        const long long base = (long long)(unit - I.unit_start) * kPlainUnit;
        float p[8], g[8], m[8], v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {   // all loads of the unit first: 32 independent requests per thread in flight
            const long long i = base + u * 256 + threadIdx.x;
            if (i < I.n) p[u] = I.p[i], g[u] = I.g[i], m[u] = I.m[i], v[u] = I.v[i];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const long long i = base + u * 256 + threadIdx.x;
            if (i < I.n) {
                adam_update(c, p[u], g[u], m[u], v[u]);
                I.p[i] = p[u], I.m[i] = m[u], I.v[i] = v[u];
            }
        }
    }
}

// ---- Adam + bf16 re-pack of convolution weights ----------------------------------------------------------------------
constexpr int kPackItems = 28;
constexpr int kPackCo = 16;
constexpr int kPackElems = 1024;   // (input channel, tap) elements staged per output channel
struct PackItem {
    float* p;
    const float* g;
    float* m;
    float* v;
    const float* step;
    __nv_bfloat16* wf;
    __nv_bfloat16* wd;
    int32_t Cout, Cin, ntaps, kpitch, copitch;
    int32_t ci_chunk, n_ci;   // input channels per block, blocks along the input channels
    int32_t unit_start;       // first block of this item
};
struct PackBatch {
    int32_t n, units;
    HyperDev h;
    PackItem it[kPackItems];
};

__global__ void __launch_bounds__(256)
adam_pack_tiled_kernel(const __grid_constant__ PackBatch B) {
    pdl_wait();
    __shared__ __align__(16) __nv_bfloat16 tile[kPackCo][kPackElems + 8];
    int lo = 0, hi = B.n - 1;   // last item with unit_start <= blockIdx.x
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (B.it[mid].unit_start <= static_cast<int>(blockIdx.x)) lo = mid;
        else hi = mid - 1;
    }
    const PackItem& I = B.it[lo];
    const AdamCoef c = adam_coef(B.h, I.step);
    const int unit = static_cast<int>(blockIdx.x) - I.unit_start;
    const int co0 = (unit / I.n_ci) * kPackCo;
    const int cc = unit % I.n_ci;
    const int ci0 = cc * I.ci_chunk;
    const int ci_len = min(I.ci_chunk, I.Cin - ci0);
    const int len = ci_len * I.ntaps;            // contiguous fp32 run of one output channel, <= kPackElems
    const int tid = threadIdx.x;
    const long long row_stride = (long long)I.Cin * I.ntaps;
    const long long org = ((long long)co0 * I.Cin + ci0) * I.ntaps;
    const int total = kPackCo * len;
    // update: rows = output channels, columns = (ci - ci0) * ntaps + tap, exactly the master weight's order.
    // Four elements per thread are loaded before any is stored (the compiler may not move loads over the stores itself).
    for (int base = tid; base < total; base += 4 * 256) {
        float p[4], g[4], m[4], v[4];
        int r[4], e[4];
        long long off[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = base + u * 256;
            r[u] = idx / len;
            e[u] = idx - r[u] * len;
            ok[u] = idx < total && co0 + r[u] < I.Cout;
            off[u] = org + r[u] * row_stride + e[u];
            if (ok[u]) p[u] = I.p[off[u]], g[u] = I.g[off[u]], m[u] = I.m[off[u]], v[u] = I.v[off[u]];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (base + u * 256 < total) {
                float w = 0.f;               // rows past Cout are staged as zeros (pad lanes of the dgrad image)
                if (ok[u]) {
                    adam_update(c, p[u], g[u], m[u], v[u]);
                    I.p[off[u]] = p[u], I.m[off[u]] = m[u], I.v[off[u]] = v[u];
                    w = p[u];
                }
                tile[r[u]][e[u]] = __float2bfloat16(w);
            }
        }
    }
    __syncthreads();
    const uint16_t* t16 = reinterpret_cast<const uint16_t*>(&tile[0][0]);
    constexpr int kRow = kPackElems + 8;
    if (I.wf != nullptr) {
        // fprop image [tap][Cout][kpitch]: this block's k range, zero beyond Cin (the last chunk also writes the pad lanes)
        const int k_end = (cc + 1 == I.n_ci) ? I.kpitch : ci0 + I.ci_chunk;
        const int kv = (k_end - ci0) >> 3;                  // 16-byte vectors per (tap, co)
        const int rows = min(kPackCo, I.Cout - co0);
        const int nvec = I.ntaps * rows * kv;
        for (int idx = tid; idx < nvec; idx += 256) {
            const int vq = idx % kv;
            const int rr = (idx / kv) % rows;
            const int tap = idx / (kv * rows);
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k0 = 8 * vq + 2 * j;              // relative to ci0
                const uint32_t lo16 = k0 < ci_len ? t16[rr * kRow + k0 * I.ntaps + tap] : 0u;
                const uint32_t hi16 = k0 + 1 < ci_len ? t16[rr * kRow + (k0 + 1) * I.ntaps + tap] : 0u;
                pk[j] = lo16 | (hi16 << 16);
            }
            __nv_bfloat16* dst = I.wf + ((long long)tap * I.Cout + co0 + rr) * I.kpitch + ci0 + 8 * vq;
            *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
    }
    if (I.wd != nullptr) {
        // dgrad image [tap][Cin][copitch]: 8 consecutive output channels per store
        const int halves = min(2, (I.copitch - co0) >> 3);
        const int nvec = I.ntaps * ci_len * halves;
        for (int idx = tid; idx < nvec; idx += 256) {
            const int hh = idx % halves;
            const int ci = (idx / halves) % ci_len;
            const int tap = idx / (halves * ci_len);
            const int col = ci * I.ntaps + tap;
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                pk[j] = static_cast<uint32_t>(t16[(8 * hh + 2 * j) * kRow + col]) |
                        (static_cast<uint32_t>(t16[(8 * hh + 2 * j + 1) * kRow + col]) << 16);
            __nv_bfloat16* dst = I.wd + ((long long)tap * I.Cin + ci0 + ci) * I.copitch + co0 + 8 * hh;
            *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
    }
}

int check_hyper(const zsv_adam_hyper* h, HyperDev* out) {
    if (!h) return fail(ZSV_ERR_BAD_ARG, "adam: null hyper-parameters");
    if (!(h->beta1 >= 0.f && h->beta1 < 1.f) || !(h->beta2 >= 0.f && h->beta2 < 1.f) || h->eps < 0.f)
        return fail(ZSV_ERR_BAD_ARG, "adam: invalid hyper-parameters");
    out->lr = h->lr, out->beta1 = h->beta1, out->beta2 = h->beta2, out->eps = h->eps, out->weight_decay = h->weight_decay;
    out->grad_scale = h->grad_scale, out->lr_dev = h->lr_dev;
    return ZSV_OK;
}

}  // namespace
}  // namespace zsv

using namespace zsv;

extern "C" int zsv_adam_step(int n, float* const* params, const float* const* grads, float* const* exp_avg,
                             float* const* exp_avg_sq, const long long* numel, const float* const* steps,
                             const zsv_adam_hyper* hyper, void* stream) {
    if (n < 0 || (n > 0 && (!params || !grads || !exp_avg || !exp_avg_sq || !numel || !steps)))
        return fail(ZSV_ERR_BAD_ARG, "adam_step: null array");
    HyperDev h;
    int rc = check_hyper(hyper, &h);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    for (int base = 0; base < n; base += kPlainItems) {
        PlainBatch B;
        memset(&B, 0, sizeof(B));
        B.h = h;
        int m = 0;
        long long units = 0;
        for (int i = base; i < std::min(n, base + kPlainItems); ++i) {
            if (numel[i] <= 0) continue;
            if (!params[i] || !grads[i] || !exp_avg[i] || !exp_avg_sq[i] || !steps[i])
                return fail(ZSV_ERR_BAD_ARG, "adam_step: null tensor %d", i);
            PlainItem& I = B.it[m++];
            I.p = params[i], I.g = grads[i], I.m = exp_avg[i], I.v = exp_avg_sq[i], I.step = steps[i], I.n = numel[i];
            I.unit_start = (int)units;
            units += ceil_div_ll(numel[i], kPlainUnit);
            if (units > 0x7fffffffLL) return fail(ZSV_ERR_UNSUPPORTED, "adam_step: too many elements in one call");
        }
        if (m == 0) continue;
        B.n = m, B.units = (int)units;
        const int blocks = (int)std::min<long long>(units, (long long)sm_count() * 8);
        zsv::launch(adam_plain_kernel, blocks, 256, 0, st, B);
        ZSV_LAUNCH_CHECK("adam_plain_kernel");
    }
    return ZSV_OK;
}

extern "C" int zsv_adam_pack_step(int n, const zsv_conv_desc* descs, float* const* params, const float* const* grads,
                                  float* const* exp_avg, float* const* exp_avg_sq, const float* const* steps,
                                  void* const* w_fprop, void* const* w_dgrad, const zsv_adam_hyper* hyper, void* stream) {
    if (n < 0 || (n > 0 && (!descs || !params || !grads || !exp_avg || !exp_avg_sq || !steps || !w_fprop || !w_dgrad)))
        return fail(ZSV_ERR_BAD_ARG, "adam_pack_step: null array");
    HyperDev h;
    int rc = check_hyper(hyper, &h);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    for (int base = 0; base < n; base += kPackItems) {
        PackBatch B;
        memset(&B, 0, sizeof(B));
        B.h = h;
        int m = 0;
        long long units = 0;
        for (int i = base; i < std::min(n, base + kPackItems); ++i) {
            const zsv_conv_desc* d = &descs[i];
            if (d->Cin < 1 || d->Cout < 1 || d->kt < 1 || d->kh < 1 || d->kw < 1)
                return fail(ZSV_ERR_BAD_ARG, "adam_pack_step: bad descriptor %d", i);
            if (d->x_layout != ZSV_CONV_X_NDHWC)
                return fail(ZSV_ERR_UNSUPPORTED, "adam_pack_step: only the plain NDHWC weight images (update the W-folded first "
                                                 "convolution with zsv_adam_step + zsv_conv3d_pack_weight)");
            const int ntaps = d->kt * d->kh * d->kw;
            if (8 * ntaps > kPackElems) return fail(ZSV_ERR_UNSUPPORTED, "adam_pack_step: more than %d taps", kPackElems / 8);
            if (!params[i] || !grads[i] || !exp_avg[i] || !exp_avg_sq[i] || !steps[i])
                return fail(ZSV_ERR_BAD_ARG, "adam_pack_step: null tensor %d", i);
            PackItem& I = B.it[m++];
            I.p = params[i], I.g = grads[i], I.m = exp_avg[i], I.v = exp_avg_sq[i], I.step = steps[i];
            I.wf = (__nv_bfloat16*)w_fprop[i], I.wd = (__nv_bfloat16*)w_dgrad[i];
            I.Cout = d->Cout, I.Cin = d->Cin, I.ntaps = ntaps, I.kpitch = cpad(d->Cin), I.copitch = cpad(d->Cout);
            I.ci_chunk = std::max(8, std::min(64, (kPackElems / ntaps) & ~7));
            I.n_ci = ceil_div(d->Cin, I.ci_chunk);
            I.unit_start = (int)units;
            // rows up to the padded channel count so that the pad lanes of the dgrad image are (re)written as zeros
            units += (long long)ceil_div(I.wd ? I.copitch : d->Cout, kPackCo) * I.n_ci;
            if (units > 0x7fffffffLL) return fail(ZSV_ERR_UNSUPPORTED, "adam_pack_step: too many tiles in one call");
        }
        if (m == 0) continue;
        B.n = m, B.units = (int)units;
        zsv::launch(adam_pack_tiled_kernel, (int)units, 256, 0, st, B);
        ZSV_LAUNCH_CHECK("adam_pack_tiled_kernel");
    }
    return ZSV_OK;
}
