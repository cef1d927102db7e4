// Implicit-GEMM 3-D convolution for sm_100a: fprop, dgrad (K-major operands) and wgrad (MN-major operands).
//
// Replaces the arithmetic PyTorch reaches from the reference's nn.Conv3d call sites
// (resnet.py:40-52 Conv2Plus1D, resnet.py:181-184 R2Plus1dStem, resnet.py:271 downsample,
// network.py:102-117 C3D) and their autograd (main.py:195).
//
// Design:
//   * activations are bf16 NDHWC; a GEMM-M tile is a (bw x bh x bt x bn) box of output positions, so one
//     5-D TMA box load at tap-shifted coordinates *is* the im2col tile: padding comes from TMA
//     out-of-bounds zero fill, stride-2 convolutions read parity-plane views of the same tensor
//     (a separate tensor map per parity, doubled strides) so every load is a plain tiled load;
//   * the TMA writes SWIZZLE_128B tiles that tcgen05.mma consumes directly from shared memory,
//     accumulating fp32 in TMEM; one elected thread issues the MMAs, one drives the TMA ring;
//   * the epilogue reads TMEM with tcgen05.ld, rounds to bf16, stores channels-last and emits
//     per-tile BatchNorm partial statistics (sum, sum of squares of the rounded values);
//   * dgrad is the same kernel on dy with the transposed weight image, decomposed into output parity
//     classes for stride 2; wgrad contracts over positions with both operands MN-major
//     (channels contiguous), split over CTAs along the position axis with deterministic partials.
#include <algorithm>
#include <mutex>
#include <stdlib.h>
#include <string.h>
#include <type_traits>
#include <vector>

#include "zsv_internal.h"
#include "zsv_ptx.cuh"

namespace zsv {
namespace {

// ------------------------------------------------------------------------------------------------
// tensor maps
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// bf16 tensor map with SWIZZLE_128B; dims/box innermost first; strides in bytes for dims 1..rank-1.
int make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
             const uint32_t* box, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(ZSV_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    cuuint64_t gd[5];
    cuuint64_t gs[4];
    cuuint32_t bx[5];
    cuuint32_t es[5];
    for (int i = 0; i < rank; ++i) {
        gd[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
    }
    for (int i = 0; i + 1 < rank; ++i) gs[i] = strides[i];
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gd, gs, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        return fail(ZSV_ERR_CUDA,
                    "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu %llu] strides [%llu %llu "
                    "%llu %llu] box [%u %u %u %u %u] base %p",
                    (int)r, rank, (unsigned long long)gd[0], (unsigned long long)(rank > 1 ? gd[1] : 0),
                    (unsigned long long)(rank > 2 ? gd[2] : 0), (unsigned long long)(rank > 3 ? gd[3] : 0),
                    (unsigned long long)(rank > 4 ? gd[4] : 0), (unsigned long long)(rank > 1 ? gs[0] : 0),
                    (unsigned long long)(rank > 2 ? gs[1] : 0), (unsigned long long)(rank > 3 ? gs[2] : 0),
                    (unsigned long long)(rank > 4 ? gs[3] : 0), bx[0], rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0,
                    rank > 3 ? bx[3] : 0, rank > 4 ? bx[4] : 0, base);
    }
    return ZSV_OK;
}

// ------------------------------------------------------------------------------------------------
// kernel argument blocks
// ------------------------------------------------------------------------------------------------
constexpr int kMaxTaps = 32;
constexpr int kMaxMaps = 8;    // parity planes of a convolution with stride 2 in all three dimensions (r3d_18)
struct alignas(64) MapPack {
    CUtensorMap m[kMaxMaps];
};
constexpr uint32_t kPanelBytes = 128 * 128;  // 128 rows x 128 B
constexpr int kWfoldWpad = 8;                // extra W columns of the ZSV_CONV_X_WFOLD layout

// Division by a runtime constant as multiply-high + shift (the divisor-specific constants come from the host): the
// persistent kernels turn a linear tile index into 4-5 box coordinates per tile in EVERY warp, and a generic 32-bit
// division is ~25 dependent instructions.  Valid for numerators below 2^31.
struct FastDiv {
    uint32_t d, mul, shr;
};
inline FastDiv make_fastdiv(int dd) {
    FastDiv f;
    f.d = dd < 1 ? 1u : (uint32_t)dd;
    f.mul = 0;
    f.shr = 0;
    if (f.d > 1) {
        uint32_t lg = 0;
        while ((1ull << lg) < f.d) ++lg;     // ceil(log2 d)
        const uint32_t p = 31 + lg;
        f.mul = (uint32_t)(((1ull << p) + f.d - 1) / f.d);
        f.shr = p - 32;
    }
    return f;
}
__device__ __forceinline__ int fdiv(int n, const FastDiv& f) {
    return f.d == 1u ? n : static_cast<int>(__umulhi(static_cast<uint32_t>(n), f.mul) >> f.shr);
}
// q = n / d, returns n % d
__device__ __forceinline__ int fdivmod(int n, const FastDiv& f, int& q) {
    q = fdiv(n, f);
    return n - q * static_cast<int>(f.d);
}

struct Tap {
    int16_t map, dw, dh, dt;  // which activation map, coordinate offsets of the box origin
    int16_t btap, r0, r1, r2; // weight tap index
};

struct IgemmArgs {
    int32_t bw, bh, bt, bn;  // box of positions forming one M tile (bw*bh*bt*bn <= 128 rows)
    int32_t tw, th, tt, tn;  // tile counts per dimension
    int32_t OW, OH, OT, ON;  // extents of the output position space
    int32_t kdim;            // reduction channels per tap (true count)
    int32_t ntaps;
    int32_t ncols;           // output columns to store (channel pitch, multiple of 8)
    int32_t nbias;           // valid entries of bias
    int32_t bn_tile;         // UMMA N
    int32_t stages;
    int32_t relu;
    int32_t tmem_cols;
    int32_t part_pitch;
    int32_t m_tiles, n_tiles;
    int32_t nstg;            // output staging buffers (2 = the TMA store of tile i overlaps the epilogue of tile i+1)
    int32_t nybuf;           // y buffers of the fused BatchNorm backward (2 = y is requested a whole tile ahead)
    FastDiv fd_ntiles, fd_tw, fd_th, fd_tt;
    int32_t debug;
    int32_t scratch_bytes;   // epilogue scratch in shared memory (BN-backward fusion sums), multiple of 1 KB
    int32_t bn_relu;
    const __nv_bfloat16* bn_y;
    const float4* bn_tab;
    float* bn_partial;       // [gridDim.x][4][ncols]: rows 0/1 = sum dz, sum dz*xhat of this CTA
    long long o_sN, o_sT, o_sH, o_sW;  // element strides of the output tensor
    __nv_bfloat16* out;
    const __nv_bfloat16* addend;
    float* part_sum;
    float* part_sq;
    const float* bias;
    Tap taps[kMaxTaps];
};

struct WgradArgs {
    int32_t bw, bh, bt, bn;  // box of positions forming one K block (rows multiple of 16, <= 128)
    int32_t tw, th, tt, tn;
    int32_t kchunks;         // 64-channel panels per tap on the x side
    int32_t npanels;         // ntaps * kchunks
    int32_t ntaps;
    int32_t ci_store;        // rows (x channels) to store per tap
    int32_t ci_pitch;        // workspace pitch (x channels)
    int32_t co_pitch;        // workspace pitch (dy channels) = n_tiles * bn_tile
    int32_t bn_tile, nbp;    // UMMA N, 64-wide dy panels per stage
    int32_t stages, tmem_cols;
    int32_t num_kb, kb_per_split;
    int32_t mt;              // M tiles (pairs of x panels) per CTA: 2 halves how often the dy panels are re-read
    int32_t acc_stride;      // TMEM columns between the accumulators of the M tiles
    float* ws;               // [split][tap][ci_pitch][co_pitch]
    Tap taps[kMaxTaps];
};

constexpr int kEpiWarps = 8;                        // 2 per TMEM lane quadrant, each takes every 2nd 16-column chunk
constexpr int kIgemmThreads = 64 + kEpiWarps * 32;

// ------------------------------------------------------------------------------------------------
// Shared epilogue of the igemm kernels (8 warps = 256 threads, `et` = thread index within them).
// One call handles one output tile: TMEM -> registers -> (+bias, +addend, ReLU) -> bf16 -> SWIZZLE_128B staging in
// shared memory -> one TMA store per 64-channel panel, plus BatchNorm partial sums read back from the staging tile.
// ------------------------------------------------------------------------------------------------
struct EpiArgs {
    const __nv_bfloat16* addend;
    const float* bias;
    float* part_sum;
    float* part_sq;
    int ncols, nbias, relu, part_pitch;
    // BatchNorm-backward fusion (dgrad): the tile written is dz = g * [ReLU mask of the BatchNorm that produced this
    // conv's input], and the CTA accumulates sum(dz) and sum(dz * y) per channel (see zsv_bn_bwd_fuse).  The y tile of
    // the output positions is TMA-loaded into shared memory (mapY, same geometry as the output map) while the tile's
    // accumulator is read out; mask and sums are then ONE pass over the two staged tiles (bn_column_pass).
    const __nv_bfloat16* bn_y;   // pre-BatchNorm tensor, same layout as the output (non-null = fusion requested)
    const float4* bn_tab;        // per channel (scale, shift, invstd, -mean*invstd); scale/shift give the ReLU mask
    int bn_relu;
    const CUtensorMap* mapY;     // tensor map of bn_y with the boxes of the output map
    uint32_t ybuf_u32;           // smem [out panels][128 rows][128 B] (SWIZZLE_128B): y of the current tile
    uint32_t bar_y, y_phase;     // "y tile landed" barrier and the parity of the current tile
    uint32_t ynext_u32, bar_ynext;   // buffer / barrier that receive the y tile of this CTA's NEXT tile
    int y_early;                 // two y buffers: the next tile's y is requested at the top of this tile's epilogue (a whole
                                 // tile ahead) instead of after this tile's column pass
    uint32_t y_rows;             // rows of one box (every 64-channel panel of a y tile brings y_rows * 128 bytes)
    int has_next, n0, n1, n2, n3, next_origin;   // box coordinates / first channel of this CTA's NEXT tile (y prefetch)
    float* st_acc;               // smem [2][ncols] running sums of this CTA: BatchNorm statistics (sum, sum of squares)
                                 // in fprop, (sum dz, sum dz*y) in the fused backward
    int remote_arrive;           // CTA pair: the TMEM-empty barrier is a shared::cluster address in the leader CTA
    int debug;   // tuning aid (ZSV_DEBUG_EPI bit mask): 1 = skip the TMA store, 2 = skip TMEM read + staging, 4 = skip stats
    uint32_t bar_full, full_phase;   // "accumulator complete" barrier of this tile: waited for inside epilogue_tile, AFTER
                                     // the global loads of the tile's first chunks are in flight
    // split epilogue (kSplit kernels): hand-over of this tile's staging buffer between the convert and the finish group
    uint32_t bar_staged, bar_free;
    uint32_t stg_phase;              // parity of this use of the staging buffer
};

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
// packed fp32 pair arithmetic (sm_100 FADD2 / FFMA2): acc += v ; acc += v*v
__device__ __forceinline__ void add2_sq2(uint64_t& s, uint64_t& q, uint32_t bf16pair) {
    uint64_t v;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "r"(bf16pair << 16), "r"(bf16pair & 0xFFFF0000u));
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(s) : "l"(v));
    asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(q) : "l"(v));
}

// shuffle stage of a transposing reduction: `n` live values per lane; lanes whose bit `m` is set keep the upper half
template <int n>
__device__ __forceinline__ void xreduce_stage(float (&v)[16], int lane, int m) {
    const bool upper = (lane & m) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
        const float send = upper ? v[i] : v[i + n / 2];
        const float keep = upper ? v[i + n / 2] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
    }
}

// BatchNorm partial statistics of one output tile: column sums and sums of squares of the bf16 values staged in shared
// memory (SWIZZLE_128B panels; rows that fall outside the tensor were staged as zeros).
// Thread = (channel octet, row segment): it walks its rows with one 16-byte load each, accumulating 8 sums and 8 sums
// of squares with packed fp32x2 adds/FMAs; the row segments of an octet sit on adjacent lanes and are combined with a
// transposing shuffle reduction, after which every lane owns one or two finished column totals and adds them to the
// CTA's running sums in shared memory (deterministic: fixed owner, fixed tile order, no atomics).
__device__ __forceinline__ void tile_column_stats(const EpiArgs& E, uint32_t staging_u32, int width, int n_origin,
                                                  int m_tile, int et, int lane) {
    const int octs = width >> 3;
    const bool seg16 = octs <= 16;                 // 16 row segments of 8 rows, else 8 segments of 16 rows
    const int segs = seg16 ? 16 : 8;
    const int ntasks = octs * segs;                // <= 256
    if ((et & ~31) >= ntasks) return;              // whole warp idle
    const bool active = et < ntasks;
    const int oct = active ? (seg16 ? et >> 4 : et >> 3) : 0;
    const int seg = et & (segs - 1);
    uint64_t s2[4] = {0ull, 0ull, 0ull, 0ull}, q2[4] = {0ull, 0ull, 0ull, 0ull};
    if (active) {
        // Segment g owns rows g, g + segs, g + 2*segs, ...: the 8 lanes of a quarter warp then read the same 16-byte
        // chunk of 8 consecutive rows, which the 128-byte swizzle spreads over all banks (conflict-free), and row & 7
        // (= g & 7) is a per-thread constant.  address(row, chunk) = row*128 + ((chunk ^ (row & 7)) << 4)
        const uint32_t a0 = staging_u32 + static_cast<uint32_t>(oct >> 3) * kPanelBytes + static_cast<uint32_t>(seg) * 128u +
                            (static_cast<uint32_t>((oct ^ seg) & 7) << 4);
        const uint32_t step = static_cast<uint32_t>(segs) * 128u;
        const int nrows = seg16 ? 8 : 16;
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (u < nrows) {
                const uint4 v = lds128(a0 + static_cast<uint32_t>(u) * step);
                add2_sq2(s2[0], q2[0], v.x);
                add2_sq2(s2[1], q2[1], v.y);
                add2_sq2(s2[2], q2[2], v.z);
                add2_sq2(s2[3], q2[3], v.w);
            }
        }
    }
    float v[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        asm("mov.b64 {%0, %1}, %2;" : "=f"(v[2 * j]), "=f"(v[2 * j + 1]) : "l"(s2[j]));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(v[8 + 2 * j]), "=f"(v[8 + 2 * j + 1]) : "l"(q2[j]));
    }
    const int colbase = n_origin + oct * 8;
    (void)m_tile;
    // Every column total has exactly one owner thread (fixed for the whole kernel), which adds it to the CTA's running
    // sum in shared memory: one partial row per CTA instead of one per tile (8624 -> 148 rows on layer 1).
    if (seg16) {
        xreduce_stage<16>(v, lane, 8);
        xreduce_stage<8>(v, lane, 4);
        xreduce_stage<4>(v, lane, 2);
        xreduce_stage<2>(v, lane, 1);
        // lane bits (of its 16-lane group): bit3 = quantity, bits 2..0 = column within the octet
        const int col = colbase + (lane & 7);
        if (active && col < E.ncols) E.st_acc[((lane & 8) ? E.ncols : 0) + col] += v[0];
    } else {
        xreduce_stage<16>(v, lane, 4);
        xreduce_stage<8>(v, lane, 2);
        xreduce_stage<4>(v, lane, 1);
        // bit2 = quantity, bit1 -> +4, bit0 -> +2, two adjacent columns per lane
        const int col = colbase + ((lane & 2) << 1) + ((lane & 1) << 1);
        if (active && col < E.ncols) {
            float* dst = E.st_acc + ((lane & 4) ? E.ncols : 0) + col;
            dst[0] += v[0];
            dst[1] += v[1];
        }
    }
}

__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Fused BatchNorm backward, per output tile: the staged tile holds g (the data gradient, bf16), ybuf the pre-BatchNorm
// values y of the same positions.  One pass, same thread layout as tile_column_stats (thread = channel octet x row
// segment, conflict-free 16-byte accesses): dz = g * [y*scale + shift > 0] is written back in place (ReLU) and the
// column sums of dz and dz*y go to the CTA's running sums.  A thread keeps one octet for the whole tile, so the eight
// (scale, shift) pairs live in registers.  Rows outside the tensor were staged as zeros and contribute nothing.
__device__ __forceinline__ void bn_column_pass(const EpiArgs& E, uint32_t staging_u32, int width, int n_origin, int et,
                                               int lane) {
    const int octs = width >> 3;
    const bool seg16 = octs <= 16;
    const int segs = seg16 ? 16 : 8;
    const int ntasks = octs * segs;
    if ((et & ~31) >= ntasks) return;              // whole warp idle
    const bool active = et < ntasks;
    const int oct = active ? (seg16 ? et >> 4 : et >> 3) : 0;
    const int seg = et & (segs - 1);
    const int colbase = n_origin + oct * 8;
    // packed fp32 pairs throughout (FFMA2 / FADD2), the ReLU mask from the packed bf16 BatchNorm output like
    // bn_bwd_reduce_kernel: ~4 instructions per channel pair and row
    f32x2 sc[4], sh[4], s2[4], q2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0;
        if (active && colbase + 2 * j < E.ncols) t0 = __ldg(E.bn_tab + colbase + 2 * j);
        if (active && colbase + 2 * j + 1 < E.ncols) t1 = __ldg(E.bn_tab + colbase + 2 * j + 1);
        sc[j] = f2_make(t0.x, t1.x), sh[j] = f2_make(t0.y, t1.y);
        s2[j] = q2[j] = 0ull;
    }
    if (active) {
        const uint32_t off0 = static_cast<uint32_t>(oct >> 3) * kPanelBytes + static_cast<uint32_t>(seg) * 128u +
                              (static_cast<uint32_t>((oct ^ seg) & 7) << 4);
        const uint32_t step = static_cast<uint32_t>(segs) * 128u;
        const int nrows = seg16 ? 8 : 16;
#pragma unroll 4
        for (int u = 0; u < nrows; ++u) {
            const uint32_t off = off0 + static_cast<uint32_t>(u) * step;
            const uint4 g = lds128(staging_u32 + off);
            const uint4 yv = lds128(E.ybuf_u32 + off);
            uint32_t gw[4] = {g.x, g.y, g.z, g.w};
            const uint32_t yw[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const f32x2 y2 = f2_from_bf16x2(yw[j]);
                if (E.bn_relu) {
                    const uint32_t bn_out = f2_to_bf16x2(f2_fma(y2, sc[j], sh[j]));
                    const __nv_bfloat162 ob = *reinterpret_cast<const __nv_bfloat162*>(&bn_out);
                    gw[j] &= __hgt2_mask(ob, __float2bfloat162_rn(0.f));
                }
                const f32x2 g2 = f2_from_bf16x2(gw[j]);
                s2[j] = f2_add(s2[j], g2);
                q2[j] = f2_fma(g2, y2, q2[j]);
            }
            if (E.bn_relu) sts128(staging_u32 + off, make_uint4(gw[0], gw[1], gw[2], gw[3]));
        }
    }
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        f2_split(s2[j], s[2 * j], s[2 * j + 1]);
        f2_split(q2[j], q[2 * j], q[2 * j + 1]);
    }
    float v[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = s[j], v[8 + j] = q[j];
    if (seg16) {
        xreduce_stage<16>(v, lane, 8);
        xreduce_stage<8>(v, lane, 4);
        xreduce_stage<4>(v, lane, 2);
        xreduce_stage<2>(v, lane, 1);
        const int col = colbase + (lane & 7);
        if (active && col < E.ncols) E.st_acc[((lane & 8) ? E.ncols : 0) + col] += v[0];
    } else {
        xreduce_stage<16>(v, lane, 4);
        xreduce_stage<8>(v, lane, 2);
        xreduce_stage<4>(v, lane, 1);
        const int col = colbase + ((lane & 2) << 1) + ((lane & 1) << 1);
        if (active && col < E.ncols) {
            float* dst = E.st_acc + ((lane & 4) ? E.ncols : 0) + col;
            dst[0] += v[0];
            dst[1] += v[1];
        }
    }
}

// (fused BatchNorm backward) start the TMA load of one y tile: panels of 64 channels, same boxes as the output stores
__device__ __forceinline__ void load_y_tile(const EpiArgs& E, uint32_t ybuf, uint32_t bar, int width, int n_origin, int o0,
                                            int o1, int o2, int o3) {
    mbar_expect_tx(bar, static_cast<uint32_t>((width + 63) >> 6) * E.y_rows * 128u);
    for (int p = 0; p * 64 < width; ++p)
        tma_load_5d(ybuf + p * kPanelBytes, E.mapY, bar, n_origin + 64 * p, o0, o1, o2, o3);
}

__device__ __forceinline__ void epilogue_tile(const EpiArgs& E, const CUtensorMap* mapOut, uint8_t* staging,
                                              uint32_t staging_u32, float* statbuf, uint32_t trow,
                                              uint32_t bar_tmem_empty, int width, int n_origin, bool valid,
                                              long long off, int o0, int o1, int o2, int o3, int m_tile,
                                              bool keep_one_store_in_flight, int q, int half, int row, int lane,
                                              int et) {
    const int srow_idx = row;           // row of the staging tile this thread writes == its TMEM lane
    // Global operands of the tile (shortcut addend, BatchNorm input of the fused backward) do not depend on the
    // accumulator: pull this thread's row into L1 before the accumulator wait, so that the loads in emit_chunk hit
    // (waiting for them was the top stall of this kernel in the ncu source view).  Prefetching into REGISTERS instead
    // pins 32 registers through the whole epilogue and was measured 20% slower on the epilogue-bound temporal convs.
    constexpr int kChunkStride = 4 * kEpiWarps;
    // (two y buffers) y of this CTA's next tile: its buffer was last read by the column pass of the previous tile
    if (et == 0 && E.bn_y != nullptr && E.has_next && E.y_early)
        load_y_tile(E, E.ynext_u32, E.bar_ynext, width, E.next_origin, E.n0, E.n1, E.n2, E.n3);
    if (E.addend != nullptr && valid && !(E.debug & 2)) {
        for (int c = 0; c < width; c += 64) {
            if (n_origin + c >= E.ncols) break;
            prefetch_l1(E.addend + off + n_origin + c);
        }
    }
    mbar_wait(E.bar_full, E.full_phase);
    tc_fence_after();
    // the staging buffer about to be written must have been read by the TMA store that used it last
    if (et == 0) {
        if (keep_one_store_in_flight) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        else tma_store_wait_read();
    }
    named_bar_sync(1, kEpiWarps * 32);
    const uint32_t srow = static_cast<uint32_t>(srow_idx) * 128u;
    const uint32_t sxor = static_cast<uint32_t>(srow_idx & 7);
    // one accumulator chunk (16 fp32 columns of this thread's row) -> bias / addend / ReLU -> bf16 -> swizzled staging
    auto emit_chunk = [&](int c, const uint32_t (&v)[16]) {
        const int col = n_origin + c;
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
        if (E.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (col + j < E.nbias) f[j] += __ldg(E.bias + col + j);
        }
        if (E.addend != nullptr && valid) {
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf) {
                if (col + 8 * hlf < E.ncols) {
                    const uint4 a = *reinterpret_cast<const uint4*>(E.addend + off + col + 8 * hlf);
                    const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        f[8 * hlf + 2 * j] += bf16_lo(aw[j]);
                        f[8 * hlf + 2 * j + 1] += bf16_hi(aw[j]);
                    }
                }
            }
        }
        if (E.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pk[j] = valid ? pack_bf16x2(f[2 * j], f[2 * j + 1]) : 0u;
        // SWIZZLE_128B staging: 16-byte chunk index XOR (row & 7) inside the 64-channel panel
        uint8_t* prow = staging + static_cast<uint32_t>(c >> 6) * kPanelBytes + srow;
        const uint32_t ch = static_cast<uint32_t>(c & 63) >> 3;
        *reinterpret_cast<uint4*>(prow + ((ch ^ sxor) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(prow + (((ch + 1) ^ sxor) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    };
    // two TMEM loads in flight per wait: the thread's chunks are 16*kEpiWarps/4 columns apart (the other warps of the
    // quadrant take the chunks in between)
    for (int c = half * 16; c < width && !(E.debug & 2); c += 2 * kChunkStride) {
        uint32_t v0[16], v1[16];
        const bool two = c + kChunkStride < width;
        tmem_ld16(trow + c, v0);
        if (two) tmem_ld16(trow + c + kChunkStride, v1);
        tmem_ld_wait();
        emit_chunk(c, v0);
        if (two) emit_chunk(c + kChunkStride, v1);
    }
    // accumulator buffer fully read: hand it back to the MMA warp
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
        if (E.remote_arrive) mbar_arrive_cluster(bar_tmem_empty);
        else mbar_arrive(bar_tmem_empty);
    }
    fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the TMA store
    named_bar_sync(2, kEpiWarps * 32);
    if (E.bn_y != nullptr) {
        // fused BatchNorm backward: ReLU mask + column sums over the staged gradient and the y tile, then the store
        mbar_wait(E.bar_y, E.y_phase);
        bn_column_pass(E, staging_u32, width, n_origin, et, lane);
        fence_proxy_async_smem();        // masked tile (generic-proxy writes) -> visible to the TMA store
        named_bar_sync(3, kEpiWarps * 32);   // every thread is done with the y tile and the staged tile is final
    }
    if (et == 0 && !(E.debug & 1)) {
        for (int p = 0; p * 64 < width; ++p) {
            const int ccol = n_origin + 64 * p;
            if (ccol < E.ncols) tma_store_5d(mapOut, staging_u32 + p * kPanelBytes, ccol, o0, o1, o2, o3);
        }
        tma_store_commit();
    }
    // (one y buffer) y of this CTA's next tile: lands while that tile's accumulator is read out
    if (et == 0 && E.bn_y != nullptr && E.has_next && !E.y_early)
        load_y_tile(E, E.ynext_u32, E.bar_ynext, width, E.next_origin, E.n0, E.n1, E.n2, E.n3);
    if (E.part_sum != nullptr && !(E.debug & 4)) tile_column_stats(E, staging_u32, width, n_origin, m_tile, et, lane);
}

// ================================================================================================
// Split epilogue (kSplit kernels; tiles of at most 64 output channels).
// The per-tile epilogue of the narrow layer-1 kernels is a serial chain that bounds them (ncu source view,
// profiles/r02_issue_loop.txt: ~1000 clk of set-up and waits, ~950 clk of TMEM read-out and ~1800 clk of statistics per
// 128 x 64 tile against 1300 clk of MMA time).  Here the eight epilogue warps form TWO groups that work on consecutive
// tiles at the same time:
//   convert group (warps 2-5, epilogue_convert): waits for the accumulator, TMEM -> registers -> (+bias, +addend, ReLU) ->
//       bf16 -> SWIZZLE_128B staging tile, hands the accumulator back to the MMA warp and the staged tile to the other group;
//   finish group (warps 6-9, epilogue_finish): fused BatchNorm-backward column pass or BatchNorm statistics over the staged
//       tile, TMA store, then frees the staging buffer ("staged" / "free" mbarriers per staging buffer; one named barrier
//       of 128 threads inside the group).
// Measured on one box (two builds): 45->64 fprop 100.8 -> 80.5 us, dgrad 71.0 -> 59.0 us, 144->64 fprop 116 -> 108 us.
// Tiles wider than 64 channels are bound by the read-out, which needs all eight warps: they keep the single group.
// ================================================================================================
constexpr int kGroupThreads = 128;

// Thread layout of the 128-thread column passes (channel octets x row segments): 16 segments of 8 rows (width <= 64).
// Segment g owns rows g, g + 16, ...: the 8 lanes of a quarter warp read one 16-byte chunk of 8 different rows, which the
// 128-byte swizzle spreads over all banks (address(row, chunk) = row*128 + ((chunk ^ (row & 7)) << 4)).
__device__ __forceinline__ void tile_column_stats_g(const EpiArgs& E, uint32_t staging_u32, int width, int n_origin, int eb,
                                                    int lane) {
    const int ntasks = (width >> 3) * 16;          // <= 128
    if ((eb & ~31) >= ntasks) return;              // whole warp idle
    const bool active = eb < ntasks;
    const int oct = active ? eb >> 4 : 0;
    const int seg = eb & 15;
    uint64_t s2[4] = {0ull, 0ull, 0ull, 0ull}, q2[4] = {0ull, 0ull, 0ull, 0ull};
    if (active) {
        const uint32_t a0 = staging_u32 + static_cast<uint32_t>(seg) * 128u + (static_cast<uint32_t>((oct ^ seg) & 7) << 4);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint4 v = lds128(a0 + static_cast<uint32_t>(u) * 2048u);
            add2_sq2(s2[0], q2[0], v.x);
            add2_sq2(s2[1], q2[1], v.y);
            add2_sq2(s2[2], q2[2], v.z);
            add2_sq2(s2[3], q2[3], v.w);
        }
    }
    float v[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        asm("mov.b64 {%0, %1}, %2;" : "=f"(v[2 * j]), "=f"(v[2 * j + 1]) : "l"(s2[j]));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(v[8 + 2 * j]), "=f"(v[8 + 2 * j + 1]) : "l"(q2[j]));
    }
    xreduce_stage<16>(v, lane, 8);
    xreduce_stage<8>(v, lane, 4);
    xreduce_stage<4>(v, lane, 2);
    xreduce_stage<2>(v, lane, 1);
    // lane bits (of its 16-lane group): bit3 = quantity, bits 2..0 = column within the octet
    const int col = n_origin + oct * 8 + (lane & 7);
    if (active && col < E.ncols) E.st_acc[((lane & 8) ? E.ncols : 0) + col] += v[0];
}

__device__ __forceinline__ void bn_column_pass_g(const EpiArgs& E, uint32_t staging_u32, int width, int n_origin, int eb,
                                                 int lane) {
    const int ntasks = (width >> 3) * 16;
    if ((eb & ~31) >= ntasks) return;
    const bool active = eb < ntasks;
    const int oct = active ? eb >> 4 : 0;
    const int seg = eb & 15;
    const int colbase = n_origin + oct * 8;
    f32x2 sc[4], sh[4], s2[4], q2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0;
        if (active && colbase + 2 * j < E.ncols) t0 = __ldg(E.bn_tab + colbase + 2 * j);
        if (active && colbase + 2 * j + 1 < E.ncols) t1 = __ldg(E.bn_tab + colbase + 2 * j + 1);
        sc[j] = f2_make(t0.x, t1.x), sh[j] = f2_make(t0.y, t1.y);
        s2[j] = q2[j] = 0ull;
    }
    if (active) {
        const uint32_t off0 = static_cast<uint32_t>(seg) * 128u + (static_cast<uint32_t>((oct ^ seg) & 7) << 4);
#pragma unroll 4
        for (int u = 0; u < 8; ++u) {
            const uint32_t off = off0 + static_cast<uint32_t>(u) * 2048u;
            const uint4 g = lds128(staging_u32 + off);
            const uint4 yv = lds128(E.ybuf_u32 + off);
            uint32_t gw[4] = {g.x, g.y, g.z, g.w};
            const uint32_t yw[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const f32x2 y2 = f2_from_bf16x2(yw[j]);
                if (E.bn_relu) {
                    const uint32_t bn_out = f2_to_bf16x2(f2_fma(y2, sc[j], sh[j]));
                    const __nv_bfloat162 ob = *reinterpret_cast<const __nv_bfloat162*>(&bn_out);
                    gw[j] &= __hgt2_mask(ob, __float2bfloat162_rn(0.f));
                }
                const f32x2 g2 = f2_from_bf16x2(gw[j]);
                s2[j] = f2_add(s2[j], g2);
                q2[j] = f2_fma(g2, y2, q2[j]);
            }
            if (E.bn_relu) sts128(staging_u32 + off, make_uint4(gw[0], gw[1], gw[2], gw[3]));
        }
    }
    float v[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        f2_split(s2[j], v[2 * j], v[2 * j + 1]);
        f2_split(q2[j], v[8 + 2 * j], v[8 + 2 * j + 1]);
    }
    xreduce_stage<16>(v, lane, 8);
    xreduce_stage<8>(v, lane, 4);
    xreduce_stage<4>(v, lane, 2);
    xreduce_stage<2>(v, lane, 1);
    const int col = colbase + (lane & 7);
    if (active && col < E.ncols) E.st_acc[((lane & 8) ? E.ncols : 0) + col] += v[0];
}

// Convert group, one tile: accumulator -> staged bf16 tile.  `row` = this thread's TMEM lane = row of the staging tile.
__device__ __forceinline__ void epilogue_convert(const EpiArgs& E, uint8_t* staging, uint32_t trow, uint32_t bar_tmem_empty,
                                                 int width, int n_origin, bool valid, long long off, int row, int lane) {
    if (E.addend != nullptr && valid && !(E.debug & 2)) {   // the shortcut addend does not depend on the accumulator
        for (int c = 0; c < width; c += 64) {
            if (n_origin + c >= E.ncols) break;
            prefetch_l1(E.addend + off + n_origin + c);
        }
    }
    mbar_wait(E.bar_full, E.full_phase);
    tc_fence_after();
    // the staging buffer about to be written: the finish group is done with its previous tile (store read, sums taken)
    mbar_wait(E.bar_free, E.stg_phase ^ 1u);
    const uint32_t srow = static_cast<uint32_t>(row) * 128u;
    const uint32_t sxor = static_cast<uint32_t>(row & 7);
    auto emit_chunk = [&](int c, const uint32_t (&v)[16]) {
        const int col = n_origin + c;
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
        if (E.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (col + j < E.nbias) f[j] += __ldg(E.bias + col + j);
        }
        if (E.addend != nullptr && valid) {
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf) {
                if (col + 8 * hlf < E.ncols) {
                    const uint4 a = *reinterpret_cast<const uint4*>(E.addend + off + col + 8 * hlf);
                    const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        f[8 * hlf + 2 * j] += bf16_lo(aw[j]);
                        f[8 * hlf + 2 * j + 1] += bf16_hi(aw[j]);
                    }
                }
            }
        }
        if (E.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pk[j] = valid ? pack_bf16x2(f[2 * j], f[2 * j + 1]) : 0u;
        uint8_t* prow = staging + static_cast<uint32_t>(c >> 6) * kPanelBytes + srow;
        const uint32_t ch = static_cast<uint32_t>(c & 63) >> 3;
        *reinterpret_cast<uint4*>(prow + ((ch ^ sxor) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(prow + (((ch + 1) ^ sxor) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    };
    // the warp reads every chunk of its lane quadrant, two TMEM loads in flight per wait
    for (int c = 0; c < width && !(E.debug & 2); c += 32) {
        uint32_t v0[16], v1[16];
        const bool two = c + 16 < width;
        tmem_ld16(trow + c, v0);
        if (two) tmem_ld16(trow + c + 16, v1);
        tmem_ld_wait();
        emit_chunk(c, v0);
        if (two) emit_chunk(c + 16, v1);
    }
    tc_fence_before();
    fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the TMA store the other group issues
    __syncwarp();
    if (lane == 0) {
        if (E.remote_arrive) mbar_arrive_cluster(bar_tmem_empty);   // accumulator buffer back to the MMA warp
        else mbar_arrive(bar_tmem_empty);
        mbar_arrive(E.bar_staged);
    }
}

// Finish group, one tile: column pass / statistics over the staged tile, TMA store, staging buffer released.
__device__ __forceinline__ void epilogue_finish(const EpiArgs& E, const CUtensorMap* mapOut, uint32_t staging_u32, int width,
                                                int n_origin, int o0, int o1, int o2, int o3, int eb, int lane) {
    // (two y buffers) y of this CTA's next tile: its buffer was last read by the column pass of the previous tile
    if (eb == 0 && E.bn_y != nullptr && E.has_next && E.y_early)
        load_y_tile(E, E.ynext_u32, E.bar_ynext, width, E.next_origin, E.n0, E.n1, E.n2, E.n3);
    mbar_wait(E.bar_staged, E.stg_phase);
    if (E.bn_y != nullptr) {
        mbar_wait(E.bar_y, E.y_phase);
        bn_column_pass_g(E, staging_u32, width, n_origin, eb, lane);
        fence_proxy_async_smem();                // masked tile (generic-proxy writes) -> visible to the TMA store
        named_bar_sync(1, kGroupThreads);        // every thread is done with the y tile and the staged tile is final
    }
    if (eb == 0 && !(E.debug & 1)) {
        for (int p = 0; p * 64 < width; ++p) {
            const int ccol = n_origin + 64 * p;
            if (ccol < E.ncols) tma_store_5d(mapOut, staging_u32 + p * kPanelBytes, ccol, o0, o1, o2, o3);
        }
        tma_store_commit();
    }
    if (eb == 0 && E.bn_y != nullptr && E.has_next && !E.y_early)
        load_y_tile(E, E.ynext_u32, E.bar_ynext, width, E.next_origin, E.n0, E.n1, E.n2, E.n3);
    if (E.part_sum != nullptr && !(E.debug & 4)) tile_column_stats_g(E, staging_u32, width, n_origin, eb, lane);
    // release the staging buffer: the store has read it and every thread of the group has taken its sums
    if (eb == 0) tma_store_wait_read();
    named_bar_sync(1, kGroupThreads);
    if (eb == 0) mbar_arrive(E.bar_free);
}

// ------------------------------------------------------------------------------------------------
// K-major implicit GEMM: out[pos][co] = sum_{tap, ci} act[pos + tap][ci] * w[tap][co][ci]
//
// Persistent, warp-specialised: grid = min(#tiles, #SMs), one CTA per SM looping over (M tile, N tile) pairs.
//   warp 0      : TMA producer -- runs ahead across tile boundaries through a `stages`-deep smem ring
//   warp 1      : MMA issuer + TMEM owner -- two accumulator buffers in TMEM, so the epilogue of tile i overlaps
//                 the mainloop of tile i+1
//   warps 2..9  : epilogue (TMEM lane quadrant = warp % 4): tcgen05.ld -> bias/addend/ReLU -> bf16 -> swizzled staging ->
//                 TMA store, plus per-tile BatchNorm partial sums or the fused BatchNorm-backward column pass.
//                 kSplit = false (tiles wider than 64 channels): one group of eight warps, the two warps of a quadrant
//                 take the 16-column chunks round-robin (the read-out is the long part; 16 warps were measured and are
//                 not faster than 8).  kSplit = true (tiles of <= 64 channels): a convert group (warps 2-5) and a finish
//                 group (warps 6-9) work on consecutive tiles at the same time, see "Split epilogue" above.
// ------------------------------------------------------------------------------------------------
// k2 = CTA-pair variant (cluster of 2, tcgen05 cta_group::2): the pair computes two M tiles (one per CTA) against the same
// N tile with M = 256 MMAs issued by the leader; every CTA loads its own A tile and HALF of the B tile (mapB then has a
// box of bn_tile/2 rows), so B-operand shared-memory reads and B TMA bytes per SM halve.  "full" barriers live in the
// leader and count both CTAs' bytes; "empty" and "TMEM full" are multicast commits; the peer's epilogue warps arrive on
// the leader's "TMEM empty" barrier through its shared::cluster address.
template <bool k2, bool kSplit>
__global__ void __launch_bounds__(kIgemmThreads, 1)
igemm_kmajor_kernel(const __grid_constant__ MapPack mapsA,
                    const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapOut,
                    const __grid_constant__ CUtensorMap mapY, const __grid_constant__ IgemmArgs P) {
    pdl_wait();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int stages = P.stages;
    const uint32_t stageB = static_cast<uint32_t>(k2 ? (P.bn_tile >> 1) : P.bn_tile) * 128u;   // B rows held by this CTA
    const uint32_t rank = k2 ? cluster_ctarank() : 0u;
    const uint32_t stageBytes = kPanelBytes + stageB;
    const uint32_t ringBytes = stages * stageBytes;
    const int out_panels = (P.bn_tile + 63) >> 6;                    // 64-channel output panels of one tile
    const uint32_t stagingOff = ringBytes;                           // [nstg][out_panels][128 rows][128 B], SWIZZLE_128B
    const uint32_t stagingBytes = out_panels * kPanelBytes;
    const uint32_t ybufOff = stagingOff + P.nstg * stagingBytes;     // y tile of the fused BatchNorm backward (one buffer)
    const uint32_t statOff = ybufOff + (P.bn_y != nullptr ? P.nybuf * stagingBytes : 0u);   // fp32 running column sums
    const uint32_t barOff = statOff + static_cast<uint32_t>(P.scratch_bytes);
    const uint32_t barFull = base + barOff;
    const uint32_t barEmpty = barFull + 8u * stages;
    const uint32_t barTmemFull = barEmpty + 8u * stages;   // [2]
    const uint32_t barTmemEmpty = barTmemFull + 16u;       // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + barOff + 16u * stages + 32u);
    const uint32_t barY = barFull + 16u * stages + 40u;   // [2]
    const uint32_t barStaged = barY + 16u;                // [2] (kSplit) staging buffer written by the 4 convert warps
    const uint32_t barFree = barStaged + 16u;             // [2] (kSplit) staging buffer released by the finish group
    constexpr int kCvtWarps = kSplit ? 4 : kEpiWarps;     // warps that read the accumulator out

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(barFull + 8u * s, 1);
            mbar_init(barEmpty + 8u * s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(barTmemFull + 8u * b, 1);
            mbar_init(barTmemEmpty + 8u * b, k2 ? 2 * kCvtWarps : kCvtWarps);
            mbar_init(barY + 8u * b, 1);
            if (kSplit) {
                mbar_init(barStaged + 8u * b, 4);
                mbar_init(barFree + 8u * b, 1);
            }
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        if (k2) tmem_alloc2(smem_u32(tmem_slot), P.tmem_cols);
        else tmem_alloc(smem_u32(tmem_slot), P.tmem_cols);
    }
    tc_fence_before();
    if (k2) cluster_sync_all();     // the peer's barriers and TMEM exist before anything of the pair touches them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t acc_stride = static_cast<uint32_t>(P.tmem_cols) >> 1;

    const int rows = P.bw * P.bh * P.bt * P.bn;
    const int kchunks = (P.kdim + 63) >> 6;
    const int num_tiles = k2 ? ((P.m_tiles + 1) >> 1) * P.n_tiles : P.m_tiles * P.n_tiles;   // pair tiles when k2
    const int first_tile = k2 ? (blockIdx.x >> 1) : blockIdx.x;
    const int tile_step = k2 ? (gridDim.x >> 1) : gridDim.x;

    // Producer and MMA warps run their loops with all 32 lanes (warp-uniform control flow and addresses); only the
    // TMA / MMA / commit instructions themselves are issued by one elected lane.  The issuing thread is a scalar
    // bottleneck (one MMA every N/2 cycles must be fed), so the loops avoid divisions, 64-bit descriptor arithmetic
    // and dynamic parameter indexing.
    if (warp == 0) {
        const uint32_t leader = elect_one();
        const uint32_t tx = (static_cast<uint32_t>(rows) * 128u + stageB) * (k2 ? 2u : 1u);   // both CTAs' bytes when k2
        const uint32_t fullLeader = k2 ? mapa_shared(barFull, 0) : barFull;
        uint32_t stage = 0, phase = 0;   // ring position and phase continue across tiles
        for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
            int m, in_;
            const int n_tile = fdivmod(tile, P.fd_ntiles, m);
            if (k2) m = 2 * m + static_cast<int>(rank);   // an M tile past the end reads zeros and stores nothing
            const int iw = fdivmod(m, P.fd_tw, m);
            const int ih = fdivmod(m, P.fd_th, m);
            const int itt = fdivmod(m, P.fd_tt, in_);
            const int w0 = iw * P.bw, h0 = ih * P.bh, t0 = itt * P.bt, n0 = in_ * P.bn;
            for (int tp = 0; tp < P.ntaps; ++tp) {
                const Tap tap = P.taps[tp];
                const CUtensorMap* mp = &mapsA.m[tap.map];
                for (int c0 = 0; c0 < P.kdim; c0 += 64) {
                    mbar_wait(barEmpty + 8u * stage, phase ^ 1u);
                    if (leader) {
                        const uint32_t sa = base + stage * stageBytes;
                        if (k2) {
                            const uint32_t full = fullLeader + 8u * stage;
                            if (rank == 0) mbar_expect_tx(barFull + 8u * stage, tx);
                            tma2_load_5d(sa, mp, full, c0, w0 + tap.dw, h0 + tap.dh, t0 + tap.dt, n0);
                            tma2_load_3d(sa + kPanelBytes, &mapB, full, c0,
                                         n_tile * P.bn_tile + static_cast<int>(rank) * (P.bn_tile >> 1), tap.btap);
                        } else {
                            const uint32_t full = barFull + 8u * stage;
                            mbar_expect_tx(full, tx);
                            tma_load_5d(sa, mp, full, c0, w0 + tap.dw, h0 + tap.dh, t0 + tap.dt, n0);
                            tma_load_3d(sa + kPanelBytes, &mapB, full, c0, n_tile * P.bn_tile, tap.btap);
                        }
                    }
                    __syncwarp();
                    if (++stage == static_cast<uint32_t>(stages)) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1 && rank == 0) {
        const uint32_t leader = elect_one();
        const uint32_t idesc = umma_idesc_bf16(k2 ? 256 : 128, P.bn_tile, 0, 0);
        const uint32_t dhi = umma_desc_hi(1024, 2);
        const int tail_steps = ((P.kdim - ((kchunks - 1) << 6)) + 15) >> 4;
        // ring position as running sums: descriptor low word of the A panel (B follows it), barrier addresses
        const uint32_t ring_lo = umma_desc_lo(base), stage16 = stageBytes >> 4, b_off16 = kPanelBytes >> 4;
        uint32_t stage = 0, phase = 0, a_lo = ring_lo, bar_full = barFull, bar_empty = barEmpty;
        uint32_t acc = 0;
        // one ring stage = one (tap, 64-channel chunk): KS k-steps of K = 16
        auto run_stage = [&](auto ks_tag, uint32_t tacc) {
            constexpr int KS = decltype(ks_tag)::value;
            mbar_wait(bar_full, phase);
            tc_fence_after();
            if (leader) {
                const uint32_t b_lo = a_lo + b_off16;
#pragma unroll
                for (int k = 0; k < KS; ++k) {
                    if (k2) umma2_bf16_lohi(tacc, a_lo + 2u * k, dhi, b_lo + 2u * k, dhi, idesc, k == 0 ? acc : 1u);
                    else umma_bf16_lohi(tacc, a_lo + 2u * k, dhi, b_lo + 2u * k, dhi, idesc, k == 0 ? acc : 1u);
                }
                if (k2) umma2_commit_mc(bar_empty, 3);
                else umma_commit(bar_empty);
            }
            acc = 1u;
            __syncwarp();
            if (++stage == static_cast<uint32_t>(stages)) {
                stage = 0, phase ^= 1u, a_lo = ring_lo, bar_full = barFull, bar_empty = barEmpty;
            } else {
                a_lo += stage16, bar_full += 8u, bar_empty += 8u;
            }
        };
        int local = 0;
        for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++local) {
            const uint32_t buf = local & 1;
            mbar_wait(barTmemEmpty + 8u * buf, ((local >> 1) & 1u) ^ 1u);   // epilogue drained this buffer
            tc_fence_after();
            const uint32_t tacc = tmem_base + buf * acc_stride;
            acc = 0;
            for (int tp = 0; tp < P.ntaps; ++tp) {
                for (int c = 0; c + 1 < kchunks; ++c) run_stage(std::integral_constant<int, 4>{}, tacc);
                if (tail_steps == 4) run_stage(std::integral_constant<int, 4>{}, tacc);
                else if (tail_steps == 1) run_stage(std::integral_constant<int, 1>{}, tacc);
                else if (tail_steps == 2) run_stage(std::integral_constant<int, 2>{}, tacc);
                else run_stage(std::integral_constant<int, 3>{}, tacc);
            }
            if (leader) {
                if (k2) umma2_commit_mc(barTmemFull + 8u * buf, 3);
                else umma_commit(barTmemFull + 8u * buf);
            }
            __syncwarp();
        }
    } else if (warp >= 2) {
      if constexpr (kSplit) {
        const bool finish = warp >= 6;        // finish group (warps 6-9) / convert group (warps 2-5)
        const int q = warp & 3;               // TMEM lane quadrant a convert warp may read
        const int row = q * 32 + lane;
        int r = row;
        const int w = r % P.bw;
        r /= P.bw;
        const int h = r % P.bh;
        r /= P.bh;
        const int t = r % P.bt;
        const int n = r / P.bt;
        const int eb = (threadIdx.x - 64) & (kGroupThreads - 1);   // thread index within the group
        EpiArgs E;
        E.addend = P.addend, E.bias = P.bias, E.part_sum = P.part_sum, E.part_sq = P.part_sq;
        E.ncols = P.ncols, E.nbias = P.nbias, E.relu = P.relu, E.part_pitch = P.part_pitch;
        E.debug = P.debug;
        E.remote_arrive = k2 ? 1 : 0;
        const uint32_t tmemEmptyBar = k2 ? mapa_shared(barTmemEmpty, 0) : barTmemEmpty;
        float* statbuf = reinterpret_cast<float*>(smem + statOff);
        E.bn_y = P.bn_y, E.bn_tab = P.bn_tab, E.bn_relu = P.bn_relu;
        E.st_acc = statbuf;
        E.mapY = &mapY, E.ybuf_u32 = base + ybufOff, E.bar_y = barY;
        E.ynext_u32 = E.ybuf_u32, E.bar_ynext = barY, E.y_early = P.nybuf == 2;
        E.y_rows = static_cast<uint32_t>(rows);
        auto tile_origin = [&](int tile, int& w0, int& h0, int& t0, int& n0, int& m_tile) {
            int m, in_;
            const int n_tile = fdivmod(tile, P.fd_ntiles, m);
            if (k2) m = 2 * m + static_cast<int>(rank);
            m_tile = m;
            const int iw = fdivmod(m, P.fd_tw, m);
            const int ih = fdivmod(m, P.fd_th, m);
            const int itt = fdivmod(m, P.fd_tt, in_);
            w0 = iw * P.bw, h0 = ih * P.bh, t0 = itt * P.bt, n0 = in_ * P.bn;
            return n_tile;
        };
        if (finish) {
            if (P.bn_y != nullptr || P.part_sum != nullptr)   // running sums of this CTA start at zero
                for (int i = eb; i < 2 * P.ncols; i += kGroupThreads) statbuf[i] = 0.f;
            if (P.bn_y != nullptr) {
                // rows of the y buffer that no box ever writes must not hold NaN bit patterns (0 * NaN in the column pass)
                uint32_t* yz = reinterpret_cast<uint32_t*>(smem + ybufOff);
                for (uint32_t i = eb; i < P.nybuf * stagingBytes / 4u; i += kGroupThreads) yz[i] = 0u;
                fence_proxy_async_smem();
            }
            named_bar_sync(1, kGroupThreads);
            if (P.bn_y != nullptr && eb == 0 && first_tile < num_tiles) {
                int w0, h0, t0, n0, mt_;
                const int nt0 = tile_origin(first_tile, w0, h0, t0, n0, mt_);
                load_y_tile(E, E.ybuf_u32, barY, P.bn_tile, nt0 * P.bn_tile, w0, h0, t0, n0);
            }
        }
        int local = 0;
        for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++local) {
            int w0, h0, t0, n0, m_tile;
            const int n_tile = tile_origin(tile, w0, h0, t0, n0, m_tile);
            const uint32_t buf = local & 1;
            // staging buffer of this tile and the parity of this use of it
            const uint32_t sbi = P.nstg == 2 ? buf : 0u;
            const uint32_t sb = sbi * stagingBytes;
            E.bar_staged = barStaged + 8u * sbi, E.bar_free = barFree + 8u * sbi;
            E.stg_phase = (P.nstg == 2 ? (local >> 1) : local) & 1u;
            if (!finish) {
                const bool valid =
                    row < rows && (w0 + w) < P.OW && (h0 + h) < P.OH && (t0 + t) < P.OT && (n0 + n) < P.ON;
                const long long off = (long long)(n0 + n) * P.o_sN + (long long)(t0 + t) * P.o_sT +
                                      (long long)(h0 + h) * P.o_sH + (long long)(w0 + w) * P.o_sW;
                E.bar_full = barTmemFull + 8u * buf, E.full_phase = (local >> 1) & 1u;
                const uint32_t trow = tmem_base + buf * acc_stride + (static_cast<uint32_t>(q * 32) << 16);
                epilogue_convert(E, smem + stagingOff + sb, trow, tmemEmptyBar + 8u * buf, P.bn_tile, n_tile * P.bn_tile,
                                 valid, off, row, lane);
            } else {
                E.has_next = tile + tile_step < num_tiles;
                if (P.bn_y != nullptr) {
                    if (E.has_next) {
                        int mt_;
                        E.next_origin = tile_origin(tile + tile_step, E.n0, E.n1, E.n2, E.n3, mt_) * P.bn_tile;
                    }
                    if (P.nybuf == 2) {   // y buffers and their barriers alternate with the tiles
                        E.ybuf_u32 = base + ybufOff + buf * stagingBytes, E.bar_y = barY + 8u * buf, E.y_phase = (local >> 1) & 1u;
                        E.ynext_u32 = base + ybufOff + (buf ^ 1u) * stagingBytes, E.bar_ynext = barY + 8u * (buf ^ 1u);
                    } else {
                        E.y_phase = local & 1u;
                    }
                }
                epilogue_finish(E, &mapOut, base + stagingOff + sb, P.bn_tile, n_tile * P.bn_tile, w0, h0, t0, n0, eb, lane);
            }
        }
        if (finish) {
            if (eb == 0) tma_store_wait_all();   // global writes of the last tile complete before the CTA exits
            if (P.bn_y != nullptr) {
                for (int i = eb; i < 2 * P.ncols; i += kGroupThreads) {   // (the last tile's named barrier ordered the sums)
                    const int qn = i >= P.ncols ? 1 : 0;
                    P.bn_partial[((long long)blockIdx.x * 4 + qn) * P.ncols + (i - qn * P.ncols)] = statbuf[i];
                }
            } else if (P.part_sum != nullptr) {
                for (int i = eb; i < 2 * P.ncols; i += kGroupThreads) {
                    const int qn = i >= P.ncols ? 1 : 0;
                    (qn ? P.part_sq : P.part_sum)[(long long)blockIdx.x * P.part_pitch + (i - qn * P.ncols)] = statbuf[i];
                }
            }
        }
      } else {
        const int q = warp & 3;               // TMEM lane quadrant this warp may read
        const int half = (warp - 2) >> 2;     // which of the warps of the quadrant (chunk index mod kEpiWarps/4)
        const int row = q * 32 + lane;
        int r = row;
        const int w = r % P.bw;
        r /= P.bw;
        const int h = r % P.bh;
        r /= P.bh;
        const int t = r % P.bt;
        const int n = r / P.bt;
        const int et = threadIdx.x - 64;      // 0..255 within the epilogue group
        EpiArgs E;
        E.addend = P.addend, E.bias = P.bias, E.part_sum = P.part_sum, E.part_sq = P.part_sq;
        E.ncols = P.ncols, E.nbias = P.nbias, E.relu = P.relu, E.part_pitch = P.part_pitch;
        E.debug = P.debug;
        E.remote_arrive = k2 ? 1 : 0;
        const uint32_t tmemEmptyBar = k2 ? mapa_shared(barTmemEmpty, 0) : barTmemEmpty;
        float* statbuf = reinterpret_cast<float*>(smem + statOff);
        E.bn_y = P.bn_y, E.bn_tab = P.bn_tab, E.bn_relu = P.bn_relu;
        E.st_acc = statbuf;
        E.mapY = &mapY, E.ybuf_u32 = base + ybufOff, E.bar_y = barY;
        E.ynext_u32 = E.ybuf_u32, E.bar_ynext = barY, E.y_early = P.nybuf == 2;
        E.y_rows = static_cast<uint32_t>(rows);
        if (P.bn_y != nullptr || P.part_sum != nullptr)   // running sums start at zero; the first tile's barriers order this before any use
            for (int i = et; i < 2 * P.ncols; i += kEpiWarps * 32) statbuf[i] = 0.f;
        // box origin / first channel of a tile of this CTA
        auto tile_origin = [&](int tile, int& w0, int& h0, int& t0, int& n0, int& m_tile) {
            int m, in_;
            const int n_tile = fdivmod(tile, P.fd_ntiles, m);
            if (k2) m = 2 * m + static_cast<int>(rank);
            m_tile = m;
            const int iw = fdivmod(m, P.fd_tw, m);
            const int ih = fdivmod(m, P.fd_th, m);
            const int itt = fdivmod(m, P.fd_tt, in_);
            w0 = iw * P.bw, h0 = ih * P.bh, t0 = itt * P.bt, n0 = in_ * P.bn;
            return n_tile;
        };
        if (P.bn_y != nullptr) {
            // rows of the y buffer that no box ever writes must not hold NaN bit patterns (0 * NaN in the column pass)
            uint32_t* yz = reinterpret_cast<uint32_t*>(smem + ybufOff);
            for (uint32_t i = et; i < P.nybuf * stagingBytes / 4u; i += kEpiWarps * 32) yz[i] = 0u;
            fence_proxy_async_smem();
            named_bar_sync(1, kEpiWarps * 32);
            if (et == 0 && first_tile < num_tiles) {
                int w0, h0, t0, n0, mt_;
                const int nt0 = tile_origin(first_tile, w0, h0, t0, n0, mt_);
                load_y_tile(E, E.ybuf_u32, barY, P.bn_tile, nt0 * P.bn_tile, w0, h0, t0, n0);
            }
        }
        int local = 0;
        for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++local) {
            int w0, h0, t0, n0, m_tile;
            const int n_tile = tile_origin(tile, w0, h0, t0, n0, m_tile);
            if (P.nybuf == 2) {   // y buffers and their barriers alternate with the tiles
                const uint32_t yb = local & 1u;
                E.ybuf_u32 = base + ybufOff + yb * stagingBytes, E.bar_y = barY + 8u * yb, E.y_phase = (local >> 1) & 1u;
                E.ynext_u32 = base + ybufOff + (yb ^ 1u) * stagingBytes, E.bar_ynext = barY + 8u * (yb ^ 1u);
            } else {
                E.y_phase = local & 1u;
            }
            E.has_next = tile + tile_step < num_tiles;
            if (P.bn_y != nullptr && E.has_next) {
                int mt_;
                E.next_origin = tile_origin(tile + tile_step, E.n0, E.n1, E.n2, E.n3, mt_) * P.bn_tile;
            }
            const bool valid =
                row < rows && (w0 + w) < P.OW && (h0 + h) < P.OH && (t0 + t) < P.OT && (n0 + n) < P.ON;
            const long long off = (long long)(n0 + n) * P.o_sN + (long long)(t0 + t) * P.o_sT +
                                  (long long)(h0 + h) * P.o_sH + (long long)(w0 + w) * P.o_sW;
            const uint32_t buf = local & 1;
            E.bar_full = barTmemFull + 8u * buf, E.full_phase = (local >> 1) & 1u;
            const uint32_t sb = (P.nstg == 2 ? (local & 1) : 0) * stagingBytes;
            const uint32_t trow = tmem_base + buf * acc_stride + (static_cast<uint32_t>(q * 32) << 16);
            epilogue_tile(E, &mapOut, smem + stagingOff + sb, base + stagingOff + sb, statbuf, trow,
                          tmemEmptyBar + 8u * buf, P.bn_tile, n_tile * P.bn_tile, valid, off, w0, h0, t0, n0, m_tile,
                          P.nstg == 2, q, half, row, lane, et);
        }
        if (et == 0) tma_store_wait_all();   // global writes of the last tile complete before the CTA exits
        if (P.bn_y != nullptr) {
            named_bar_sync(1, kEpiWarps * 32);   // every column owner has added its last tile
            for (int i = et; i < 2 * P.ncols; i += kEpiWarps * 32) {
                const int qn = i >= P.ncols ? 1 : 0;
                P.bn_partial[((long long)blockIdx.x * 4 + qn) * P.ncols + (i - qn * P.ncols)] = statbuf[i];
            }
        } else if (P.part_sum != nullptr) {
            named_bar_sync(1, kEpiWarps * 32);
            for (int i = et; i < 2 * P.ncols; i += kEpiWarps * 32) {
                const int qn = i >= P.ncols ? 1 : 0;
                (qn ? P.part_sq : P.part_sum)[(long long)blockIdx.x * P.part_pitch + (i - qn * P.ncols)] = statbuf[i];
            }
        }
    
      }
    }
    if (k2) cluster_sync_all();     // neither CTA leaves (or frees TMEM) while the pair still reads its shared memory
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (k2) tmem_dealloc2(tmem_base, P.tmem_cols);
        else tmem_dealloc(tmem_base, P.tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------------
// Halo implicit GEMM for stride-1 convolutions whose weight image fits in shared memory
// (spatial 1x3x3 and temporal 3x1x1 of the stem / layer 1; fprop and dgrad).
//
// TMA ingest per SM (~40 B/clk) is what bounds the generic kernel: one load per tap brings ~34 MAC per byte.  Here
//   * the weight tile of this CTA (all taps, all channel chunks) is loaded ONCE and stays resident, and
//   * the activation box is ordered with the tap ("shift") dimension OUTERMOST in shared memory and carries a halo of
//     S-1 extra slices, so the S taps along that dimension are the same staged tile read at descriptor offsets of
//     whole swizzle atoms (inner_rows % 8 == 0);
//   * taps along W (spatial convs): with b[0] == 8 the box is widened by kw-1 columns and the W taps are start offsets
//     of one 128-byte row each -- every 8-row group of the A operand is one W run, consecutive groups are w_ext rows
//     apart (descriptor SBO = w_ext * 128); the swizzle is a function of the absolute shared-memory address, so start
//     addresses and group strides that are not multiples of the 1 KB atom address exactly the rows TMA wrote
//     (bit-exact against the copy version, tests/test_gpu_conv.py).  Other box widths load kw W-shifted copies.
// One staged chunk therefore feeds kw x S x ksteps MMAs: ~420 MAC per ingested byte for the 64->144 conv.
// Channel tails of 16 / 32 use SWIZZLE_32B / SWIZZLE_64B tiles so they cost 1/4 / 1/2 of a full chunk.
// Same persistent warp-specialised structure and epilogue (single group or kSplit) as igemm_kmajor_kernel; a CTA keeps
// one N tile.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxCopies = 8;

// MMAs of one ring stage of the halo kernel: `ncp` W taps x 3 taps along the shift dimension x KS k-steps (K = 16
// each), fully unrolled per W tap.  Descriptor low words advance by running sums (a_tap / b_tap per shift tap, a_cp / b_cp
// per W tap; unsigned wrap-around encodes negative steps), the high words never change.
template <bool k2, int KS>
__device__ __forceinline__ void issue_stage3(uint32_t tacc, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t a_tap, uint32_t b_tap, uint32_t a_cp, uint32_t b_cp, int ncp,
                                             uint32_t idesc, uint32_t acc) {
    for (int cp = 0; cp < ncp; ++cp) {
        uint32_t a = a_lo, b = b_lo;
        if (k2) {
#pragma unroll
            for (int sh = 0; sh < 3; ++sh) {
#pragma unroll
                for (int k = 0; k < KS; ++k)
                    umma2_bf16_lohi(tacc, a + 2u * k, a_hi, b + 2u * k, b_hi, idesc, (sh == 0 && k == 0) ? acc : 1u);
                a += a_tap;
                b += b_tap;
            }
        } else {
            // single CTA (the temporal convolutions): measured faster with the shift taps as a loop (one box, A/B of
            // two builds, profiles/r02_issue_loop.txt: 144->64 dgrad 102 us against 112 us fully unrolled)
#pragma unroll 1
            for (int sh = 0; sh < 3; ++sh) {
#pragma unroll
                for (int k = 0; k < KS; ++k) umma_bf16_lohi(tacc, a + 2u * k, a_hi, b + 2u * k, b_hi, idesc, k == 0 ? acc : 1u);
                acc = 1u;
                a += a_tap;
                b += b_tap;
            }
        }
        acc = 1u;
        a_lo += a_cp;
        b_lo += b_cp;
    }
}


struct HaloArgs {
    int32_t b[4];        // box extents in smem row order: inner dims 0..2, then the shift dim (output extent)
    int32_t tl[4];       // tile counts per box dim
    int32_t O[4];        // output extents per box dim
    long long os[4];     // output element strides per box dim (addend addressing)
    int32_t S;           // taps along the shift dimension
    int32_t ncopies;     // W-shifted copies per channel chunk
    int32_t copy_off[kMaxCopies];   // coordinate offset along box dim 0 for each copy
    int32_t shift_org;   // coordinate of the halo origin relative to the tile origin along the shift dim
    int32_t tap0, tap_dcp, tap_dsh; // weight tap index for (copy cp, shift tap sh) = tap0 + cp*tap_dcp + sh*tap_dsh
    int32_t kdim, nchunks, tail_box; // reduction channels, 64-wide chunks (incl. tail), box width of the tail chunk
    int32_t ntaps;
    int32_t ncols, nbias, bn_tile, n_step, n_tiles, m_tiles, stages, relu, tmem_cols, part_pitch, nstg;
    uint32_t a_stage_bytes, b_main_bytes, b_tail_bytes, b_total_bytes;
    FastDiv fd_tl0, fd_tl1, fd_tl2;
    int32_t debug;
    int32_t pf_dist;     // L2 prefetch distance in tiles (0 = off)
    int32_t scratch_bytes;   // epilogue scratch in shared memory (BN-backward fusion sums), multiple of 1 KB
    int32_t nybuf;           // y buffers of the fused BatchNorm backward (2 = y is requested a whole tile ahead)
    int32_t wshift;          // 1: ONE box per chunk, widened by kw-1 along W; the W taps are descriptor start offsets
    int32_t w_ext;           // box extent along W in that mode (b[0] + kw - 1)
    int32_t min_off;         // smallest copy_off: W origin of the widened box relative to the tile
    int32_t w_first, w_step; // row of W tap cp inside the widened box = w_first + cp * w_step (copy_off is arithmetic)
    int32_t bn_relu;
    const __nv_bfloat16* bn_y;
    const float4* bn_tab;
    float* bn_partial;       // [gridDim.x][4][ncols]
    const __nv_bfloat16* addend;
    float* part_sum;
    float* part_sq;
    const float* bias;
};

// k2 = true: CTA pair (cluster of 2, tcgen05 cta_group::2) on M = 256 tiles, each CTA holding HALF of the resident weight
// rows -- the shared memory that frees is what buys a deep activation ring for the 9-tap 64<->144 convolutions, whose
// 162 KB weight image otherwise leaves two ring stages (every load latency exposed).  Barrier protocol as in
// igemm_kmajor_kernel<true, *>.
template <bool k2, bool kSplit>
__global__ void __launch_bounds__(kIgemmThreads, 1)
igemm_halo_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapAtail,
                  const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapBtail,
                  const __grid_constant__ CUtensorMap mapOut, const __grid_constant__ CUtensorMap mapY,
                  const __grid_constant__ HaloArgs P) {
    pdl_wait();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int stages = P.stages;
    const uint32_t rank = k2 ? cluster_ctarank() : 0u;
    const int b_rows = k2 ? (P.bn_tile >> 1) : P.bn_tile;   // weight rows resident in THIS CTA
    const uint32_t ringOff = P.b_total_bytes;
    const uint32_t ringBytes = stages * P.a_stage_bytes;
    const int out_panels = (P.bn_tile + 63) >> 6;
    const uint32_t stagingOff = ringOff + ringBytes;
    const uint32_t stagingBytes = out_panels * kPanelBytes;
    const uint32_t ybufOff = stagingOff + P.nstg * stagingBytes;     // y tile of the fused BatchNorm backward (one buffer)
    const uint32_t statOff = ybufOff + (P.bn_y != nullptr ? P.nybuf * stagingBytes : 0u);
    const uint32_t barOff = statOff + static_cast<uint32_t>(P.scratch_bytes);
    const uint32_t barFull = base + barOff;
    const uint32_t barEmpty = barFull + 8u * stages;
    const uint32_t barTmemFull = barEmpty + 8u * stages;
    const uint32_t barTmemEmpty = barTmemFull + 16u;
    const uint32_t barB = barTmemEmpty + 16u;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + barOff + 16u * stages + 40u);
    const uint32_t barY = barFull + 16u * stages + 48u;   // [2]
    const uint32_t barStaged = barY + 16u;                // [2] (kSplit) staging buffer written by the 4 convert warps
    const uint32_t barFree = barStaged + 16u;             // [2] (kSplit) staging buffer released by the finish group
    constexpr int kCvtWarps = kSplit ? 4 : kEpiWarps;     // warps that read the accumulator out

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(barFull + 8u * s, 1);
            mbar_init(barEmpty + 8u * s, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(barTmemFull + 8u * i, 1);
            mbar_init(barTmemEmpty + 8u * i, k2 ? 2 * kCvtWarps : kCvtWarps);
            mbar_init(barY + 8u * i, 1);
            if (kSplit) {
                mbar_init(barStaged + 8u * i, 4);
                mbar_init(barFree + 8u * i, 1);
            }
        }
        mbar_init(barB, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        if (k2) tmem_alloc2(smem_u32(tmem_slot), P.tmem_cols);
        else tmem_alloc(smem_u32(tmem_slot), P.tmem_cols);
    }
    tc_fence_before();
    if (k2) cluster_sync_all();
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t acc_stride = static_cast<uint32_t>(P.tmem_cols) >> 1;

    const int inner_rows = P.b[0] * P.b[1] * P.b[2];
    const int rows = inner_rows * P.b[3];
    // rows of one staged box: with wshift every W run of b[0] positions carries its kw-1 halo columns
    const int ld_inner_rows = P.wshift ? P.w_ext * P.b[1] * P.b[2] : inner_rows;
    const int halo_rows = ld_inner_rows * (P.b[3] + P.S - 1);
    const int nload = P.wshift ? 1 : P.ncopies;   // TMA loads (ring stages) per channel chunk
    const int nmain = (P.tail_box == 64) ? P.nchunks : P.nchunks - 1;   // chunks that use the 128-byte-row maps
    const uint32_t tail_row_bytes = static_cast<uint32_t>(P.tail_box) * 2u;
    const uint32_t per_tap_bytes = nmain * P.b_main_bytes + (nmain < P.nchunks ? P.b_tail_bytes : 0u);
    // this CTA's N tile is fixed (its weight tile is resident); M tiles are strided over the CTAs that share it
    // (a CTA pair takes M tiles 2i and 2i+1 and always the whole N range)
    const int n_tile = k2 ? 0 : blockIdx.x % P.n_tiles;
    const int m_first = k2 ? (blockIdx.x >> 1) : blockIdx.x / P.n_tiles;
    const int m_stride = k2 ? (gridDim.x >> 1) : gridDim.x / P.n_tiles;
    const int m_count = k2 ? ((P.m_tiles + 1) >> 1) : P.m_tiles;
    const int n_origin = n_tile * P.n_step;

    if (warp == 0) {
        const uint32_t leader = elect_one();
        if (leader) {
            // resident weights: [tap][chunk] tiles
            uint32_t btx = 0;
            for (int c = 0; c < P.nchunks; ++c) btx += static_cast<uint32_t>(b_rows) * (c < nmain ? 128u : tail_row_bytes);
            const uint32_t barBLeader = k2 ? mapa_shared(barB, 0) : barB;   // the pair's MMA issuer waits for both halves
            if (rank == 0) mbar_expect_tx(barB, btx * P.ntaps * (k2 ? 2u : 1u));
            for (int tp = 0; tp < P.ntaps; ++tp)
                for (int c = 0; c < P.nchunks; ++c) {
                    const uint32_t dst = base + tp * per_tap_bytes + (c < nmain ? c * P.b_main_bytes : nmain * P.b_main_bytes);
                    if (k2) tma2_load_3d(dst, c < nmain ? &mapB : &mapBtail, barBLeader, c << 6,
                                         n_origin + static_cast<int>(rank) * b_rows, tp);
                    else tma_load_3d(dst, c < nmain ? &mapB : &mapBtail, barB, c << 6, n_origin, tp);
                }
        }
        __syncwarp();
        const uint32_t fullLeader = k2 ? mapa_shared(barFull, 0) : barFull;
        uint32_t stage = 0, phase = 0;
        for (int it = m_first; it < m_count; it += m_stride) {
            const int mt = k2 ? 2 * it + static_cast<int>(rank) : it;   // an M tile past the end reads zeros, stores nothing
            int m, m3;
            const int o0 = fdivmod(mt, P.fd_tl0, m) * P.b[0];
            const int o1 = fdivmod(m, P.fd_tl1, m) * P.b[1];
            const int o2 = fdivmod(m, P.fd_tl2, m3) * P.b[2];
            const int o3 = m3 * P.b[3];
            for (int c = 0; c < P.nchunks; ++c) {
                const bool main_chunk = c < nmain;
                const uint32_t tx = static_cast<uint32_t>(halo_rows) * (main_chunk ? 128u : tail_row_bytes);
                const CUtensorMap* mp = main_chunk ? &mapA : &mapAtail;
                for (int cp = 0; cp < nload; ++cp) {
                    const int w_org = o0 + (P.wshift ? P.min_off : P.copy_off[cp]);
                    mbar_wait(barEmpty + 8u * stage, phase ^ 1u);
                    if (leader) {
                        if (k2) {
                            if (rank == 0) mbar_expect_tx(barFull + 8u * stage, 2u * tx);
                            tma2_load_5d(base + ringOff + stage * P.a_stage_bytes, mp, fullLeader + 8u * stage, c << 6,
                                         w_org, o1, o2, o3 + P.shift_org);
                        } else {
                            const uint32_t full = barFull + 8u * stage;
                            mbar_expect_tx(full, tx);
                            tma_load_5d(base + ringOff + stage * P.a_stage_bytes, mp, full, c << 6, w_org, o1, o2,
                                        o3 + P.shift_org);
                        }
                        // Pull the same box of a tile `pf_dist` iterations ahead into L2: a TMA load keeps one request
                        // per 128-byte row outstanding until its data returns, which caps DRAM-sourced loads near
                        // 3 TB/s chip-wide; L2 hits return ~6x sooner.
                        if (cp == 0 && P.pf_dist > 0) {
                            const int mp_ = mt + P.pf_dist * m_stride * (k2 ? 2 : 1);
                            if (mp_ < P.m_tiles) {
                                int q1, q2, q3;
                                const int p0 = fdivmod(mp_, P.fd_tl0, q1) * P.b[0];
                                const int p1 = fdivmod(q1, P.fd_tl1, q2) * P.b[1];
                                const int p2 = fdivmod(q2, P.fd_tl2, q3) * P.b[2];
                                tma_prefetch_5d(mp, c << 6, p0 + (P.wshift ? P.min_off : P.copy_off[0]), p1, p2,
                                                q3 * P.b[3] + P.shift_org);
                            }
                        }
                    }
                    __syncwarp();
                    if (++stage == static_cast<uint32_t>(stages)) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // The issuing thread is a scalar bottleneck of the small-N convolutions (ncu source view, profiles/r02_issue_loop.txt:
        // it never waited, it ran 16-20 instructions per MMA at 5-7 clk each while the tensor pipe idled), so everything
        // that does not change from stage to stage is computed once, descriptor words and barrier addresses are running
        // sums, and a stage is issued as straight-line code (issue_stage3).
        const uint32_t leader = elect_one();
        const uint32_t idesc = umma_idesc_bf16(k2 ? 256 : 128, P.bn_tile, 0, 0);
        const uint32_t tail_layout = P.tail_box == 16 ? 6u : (P.tail_box == 32 ? 4u : 2u);
        const int tail_steps = ((P.kdim - ((P.nchunks - 1) << 6)) + 15) >> 4;
        const uint32_t tap_bytes16 = per_tap_bytes >> 4;   // descriptor start addresses count 16-byte units
        // the chunks before the last are 64 channels wide (128-byte rows, four k-steps); the last one may use the
        // narrow-row maps (tail_box 16 / 32) and fewer k-steps
        const bool narrow_last = nmain < P.nchunks;
        const uint32_t rb_last = narrow_last ? tail_row_bytes : 128u;
        const uint32_t bhi_main = umma_desc_hi(1024, 2);
        const uint32_t bhi_last = narrow_last ? umma_desc_hi(tail_row_bytes * 8u, tail_layout) : bhi_main;
        // wshift: consecutive 8-row groups of the A operand (one W run each) are w_ext rows apart in the staged box
        const uint32_t ahi_main = P.wshift ? umma_desc_hi(static_cast<uint32_t>(P.w_ext) * 128u, 2) : bhi_main;
        const uint32_t ahi_last = P.wshift ? umma_desc_hi(static_cast<uint32_t>(P.w_ext) * rb_last, narrow_last ? tail_layout : 2u)
                                           : bhi_last;
        const uint32_t tap16_main = (static_cast<uint32_t>(ld_inner_rows) * 128u) >> 4;     // one slice along the shift dim
        const uint32_t tap16_last = (static_cast<uint32_t>(ld_inner_rows) * rb_last) >> 4;
        const uint32_t w0_main = P.wshift ? (static_cast<uint32_t>(P.w_first) * 128u) >> 4 : 0u;   // row of the first W tap
        const uint32_t w0_last = P.wshift ? (static_cast<uint32_t>(P.w_first) * rb_last) >> 4 : 0u;
        const uint32_t cp16_main = P.wshift ? (static_cast<uint32_t>(P.w_step) * 128u) >> 4 : 0u;  // from W tap to W tap
        const uint32_t cp16_last = P.wshift ? (static_cast<uint32_t>(P.w_step) * rb_last) >> 4 : 0u;
        const uint32_t b_cp = static_cast<uint32_t>(P.tap_dcp) * tap_bytes16;     // may wrap (negative step)
        const uint32_t b_step = static_cast<uint32_t>(P.tap_dsh) * tap_bytes16;   // may wrap (negative step)
        const uint32_t b_first = umma_desc_lo(base) + static_cast<uint32_t>(P.tap0) * tap_bytes16;   // chunk 0, first tap
        const uint32_t b_chunk16 = P.b_main_bytes >> 4;
        const uint32_t ring_lo = umma_desc_lo(base + ringOff), stage16 = P.a_stage_bytes >> 4;
        const int ncp = P.wshift ? P.ncopies : 1;   // W taps per stage
        const int full_chunks = P.nchunks - 1;
        mbar_wait(barB, 0);
        tc_fence_after();
        uint32_t stage = 0, phase = 0, stage_lo = ring_lo, bar_full = barFull, bar_empty = barEmpty;
        uint32_t acc = 0;
        // one chunk = `nload` ring stages (one per W copy, or a single widened box)
        auto run_chunk = [&](auto ks_tag, uint32_t tacc, uint32_t w0, uint32_t ahi, uint32_t b_lo, uint32_t bhi, uint32_t tap16,
                             uint32_t cp16) {
            constexpr int KS = decltype(ks_tag)::value;
            for (int ld = 0; ld < nload; ++ld) {
                mbar_wait(bar_full, phase);
                tc_fence_after();
                if (P.debug & 8) {   // tuning aid: consume the stage without issuing MMAs
                    if (leader) {
                        mbar_arrive(bar_empty);
                        if (k2) mbar_arrive_cluster(mapa_shared(bar_empty, 1));
                    }
                } else if (leader) {
                    issue_stage3<k2, KS>(tacc, stage_lo + w0, ahi, b_lo, bhi, tap16, b_step, cp16, b_cp, ncp, idesc, acc);
                    if (k2) umma2_commit_mc(bar_empty, 3);
                    else umma_commit(bar_empty);
                }
                acc = 1u;
                __syncwarp();
                b_lo += b_cp;   // copies mode: stage `ld` carries W tap `ld`
                if (++stage == static_cast<uint32_t>(stages)) {
                    stage = 0, phase ^= 1u, stage_lo = ring_lo, bar_full = barFull, bar_empty = barEmpty;
                } else {
                    stage_lo += stage16, bar_full += 8u, bar_empty += 8u;
                }
            }
        };
        int local = 0;
        for (int it = m_first; it < m_count; it += m_stride, ++local) {
            const uint32_t buf = local & 1;
            mbar_wait(barTmemEmpty + 8u * buf, ((local >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t tacc = tmem_base + buf * acc_stride;
            acc = 0;   // 0 only for the very first MMA of the tile
            uint32_t b_lo = b_first;
            for (int c = 0; c < full_chunks; ++c, b_lo += b_chunk16)
                run_chunk(std::integral_constant<int, 4>{}, tacc, w0_main, ahi_main, b_lo, bhi_main, tap16_main, cp16_main);
            if (tail_steps == 4) run_chunk(std::integral_constant<int, 4>{}, tacc, w0_last, ahi_last, b_lo, bhi_last, tap16_last, cp16_last);
            else if (tail_steps == 1) run_chunk(std::integral_constant<int, 1>{}, tacc, w0_last, ahi_last, b_lo, bhi_last, tap16_last, cp16_last);
            else if (tail_steps == 2) run_chunk(std::integral_constant<int, 2>{}, tacc, w0_last, ahi_last, b_lo, bhi_last, tap16_last, cp16_last);
            else run_chunk(std::integral_constant<int, 3>{}, tacc, w0_last, ahi_last, b_lo, bhi_last, tap16_last, cp16_last);
            if (leader) {
                if (P.debug & 8) {
                    mbar_arrive(barTmemFull + 8u * buf);
                    if (k2) mbar_arrive_cluster(mapa_shared(barTmemFull + 8u * buf, 1));
                } else if (k2) {
                    umma2_commit_mc(barTmemFull + 8u * buf, 3);
                } else {
                    umma_commit(barTmemFull + 8u * buf);
                }
            }
            __syncwarp();
        }
    } else if (warp >= 2) {
      if constexpr (kSplit) {
        const bool finish = warp >= 6;        // finish group (warps 6-9) / convert group (warps 2-5)
        const int q = warp & 3;
        const int row = q * 32 + lane;
        int r = row;
        const int i0 = r % P.b[0];
        r /= P.b[0];
        const int i1 = r % P.b[1];
        r /= P.b[1];
        const int i2 = r % P.b[2];
        const int i3 = r / P.b[2];
        const int eb = (threadIdx.x - 64) & (kGroupThreads - 1);   // thread index within the group
        EpiArgs E;
        E.addend = P.addend, E.bias = P.bias, E.part_sum = P.part_sum, E.part_sq = P.part_sq;
        E.ncols = P.ncols, E.nbias = P.nbias, E.relu = P.relu, E.part_pitch = P.part_pitch;
        E.debug = P.debug;
        E.remote_arrive = k2 ? 1 : 0;
        const uint32_t tmemEmptyBar = k2 ? mapa_shared(barTmemEmpty, 0) : barTmemEmpty;
        float* statbuf = reinterpret_cast<float*>(smem + statOff);
        E.bn_y = P.bn_y, E.bn_tab = P.bn_tab, E.bn_relu = P.bn_relu;
        E.st_acc = statbuf;
        E.mapY = &mapY, E.ybuf_u32 = base + ybufOff, E.bar_y = barY;
        E.ynext_u32 = E.ybuf_u32, E.bar_ynext = barY, E.y_early = P.nybuf == 2;
        E.y_rows = static_cast<uint32_t>(rows);
        // columns this tile owns: a non-last N tile only owns n_step of its bn_tile computed columns
        const int width = (n_tile + 1 < P.n_tiles) ? P.n_step : P.bn_tile;
        auto tile_origin = [&](int it, int& o0, int& o1, int& o2, int& o3) {
            const int mt = k2 ? 2 * it + static_cast<int>(rank) : it;
            int m, m3;
            o0 = fdivmod(mt, P.fd_tl0, m) * P.b[0];
            o1 = fdivmod(m, P.fd_tl1, m) * P.b[1];
            o2 = fdivmod(m, P.fd_tl2, m3) * P.b[2];
            o3 = m3 * P.b[3];
            return mt;
        };
        if (finish) {
            if (P.bn_y != nullptr || P.part_sum != nullptr)   // running sums of this CTA start at zero
                for (int i = eb; i < 2 * P.ncols; i += kGroupThreads) statbuf[i] = 0.f;
            if (P.bn_y != nullptr) {
                // rows of the y buffer that no box ever writes must not hold NaN bit patterns (0 * NaN in the column pass)
                uint32_t* yz = reinterpret_cast<uint32_t*>(smem + ybufOff);
                for (uint32_t i = eb; i < P.nybuf * stagingBytes / 4u; i += kGroupThreads) yz[i] = 0u;
                fence_proxy_async_smem();
            }
            named_bar_sync(1, kGroupThreads);
            if (P.bn_y != nullptr && eb == 0 && m_first < m_count) {
                int o0, o1, o2, o3;
                tile_origin(m_first, o0, o1, o2, o3);
                load_y_tile(E, E.ybuf_u32, barY, width, n_origin, o0, o1, o2, o3);
            }
        }
        E.next_origin = n_origin;
        int local = 0;
        for (int it = m_first; it < m_count; it += m_stride, ++local) {
            int o0, o1, o2, o3;
            tile_origin(it, o0, o1, o2, o3);
            const uint32_t buf = local & 1;
            // staging buffer of this tile and the parity of this use of it
            const uint32_t sbi = P.nstg == 2 ? buf : 0u;
            const uint32_t sb = sbi * stagingBytes;
            E.bar_staged = barStaged + 8u * sbi, E.bar_free = barFree + 8u * sbi;
            E.stg_phase = (P.nstg == 2 ? (local >> 1) : local) & 1u;
            if (!finish) {
                const bool valid = row < rows && (o0 + i0) < P.O[0] && (o1 + i1) < P.O[1] && (o2 + i2) < P.O[2] &&
                                   (o3 + i3) < P.O[3];
                const long long off = (long long)(o0 + i0) * P.os[0] + (long long)(o1 + i1) * P.os[1] +
                                      (long long)(o2 + i2) * P.os[2] + (long long)(o3 + i3) * P.os[3];
                E.bar_full = barTmemFull + 8u * buf, E.full_phase = (local >> 1) & 1u;
                const uint32_t trow = tmem_base + buf * acc_stride + (static_cast<uint32_t>(q * 32) << 16);
                epilogue_convert(E, smem + stagingOff + sb, trow, tmemEmptyBar + 8u * buf, width, n_origin, valid, off, row,
                                 lane);
            } else {
                E.has_next = it + m_stride < m_count;
                if (P.bn_y != nullptr) {
                    if (E.has_next) tile_origin(it + m_stride, E.n0, E.n1, E.n2, E.n3);
                    if (P.nybuf == 2) {   // y buffers and their barriers alternate with the tiles
                        E.ybuf_u32 = base + ybufOff + buf * stagingBytes, E.bar_y = barY + 8u * buf, E.y_phase = (local >> 1) & 1u;
                        E.ynext_u32 = base + ybufOff + (buf ^ 1u) * stagingBytes, E.bar_ynext = barY + 8u * (buf ^ 1u);
                    } else {
                        E.y_phase = local & 1u;
                    }
                }
                epilogue_finish(E, &mapOut, base + stagingOff + sb, width, n_origin, o0, o1, o2, o3, eb, lane);
            }
        }
        if (finish) {
            if (eb == 0) tma_store_wait_all();
            if (P.bn_y != nullptr) {
                for (int i = eb; i < 2 * P.ncols; i += kGroupThreads) {   // (the last tile's named barrier ordered the sums)
                    const int qn = i >= P.ncols ? 1 : 0;
                    P.bn_partial[((long long)blockIdx.x * 4 + qn) * P.ncols + (i - qn * P.ncols)] = statbuf[i];
                }
            } else if (P.part_sum != nullptr) {
                for (int i = eb; i < 2 * P.ncols; i += kGroupThreads) {
                    const int qn = i >= P.ncols ? 1 : 0;
                    (qn ? P.part_sq : P.part_sum)[(long long)blockIdx.x * P.part_pitch + (i - qn * P.ncols)] = statbuf[i];
                }
            }
        }
      } else {
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        int r = row;
        const int i0 = r % P.b[0];
        r /= P.b[0];
        const int i1 = r % P.b[1];
        r /= P.b[1];
        const int i2 = r % P.b[2];
        const int i3 = r / P.b[2];
        const int et = threadIdx.x - 64;
        EpiArgs E;
        E.addend = P.addend, E.bias = P.bias, E.part_sum = P.part_sum, E.part_sq = P.part_sq;
        E.ncols = P.ncols, E.nbias = P.nbias, E.relu = P.relu, E.part_pitch = P.part_pitch;
        E.debug = P.debug;
        E.remote_arrive = k2 ? 1 : 0;
        const uint32_t tmemEmptyBar = k2 ? mapa_shared(barTmemEmpty, 0) : barTmemEmpty;
        float* statbuf = reinterpret_cast<float*>(smem + statOff);
        E.bn_y = P.bn_y, E.bn_tab = P.bn_tab, E.bn_relu = P.bn_relu;
        E.st_acc = statbuf;
        E.mapY = &mapY, E.ybuf_u32 = base + ybufOff, E.bar_y = barY;
        E.ynext_u32 = E.ybuf_u32, E.bar_ynext = barY, E.y_early = P.nybuf == 2;
        E.y_rows = static_cast<uint32_t>(rows);
        if (P.bn_y != nullptr || P.part_sum != nullptr)   // running sums start at zero; the first tile's barriers order this before any use
            for (int i = et; i < 2 * P.ncols; i += kEpiWarps * 32) statbuf[i] = 0.f;
        // columns this tile owns: a non-last N tile only owns n_step of its bn_tile computed columns
        const int width = (n_tile + 1 < P.n_tiles) ? P.n_step : P.bn_tile;
        // box origin of M tile `it` of this CTA
        auto tile_origin = [&](int it, int& o0, int& o1, int& o2, int& o3) {
            const int mt = k2 ? 2 * it + static_cast<int>(rank) : it;
            int m, m3;
            o0 = fdivmod(mt, P.fd_tl0, m) * P.b[0];
            o1 = fdivmod(m, P.fd_tl1, m) * P.b[1];
            o2 = fdivmod(m, P.fd_tl2, m3) * P.b[2];
            o3 = m3 * P.b[3];
            return mt;
        };
        if (P.bn_y != nullptr) {
            // rows of the y buffer that no box ever writes must not hold NaN bit patterns (0 * NaN in the column pass)
            uint32_t* yz = reinterpret_cast<uint32_t*>(smem + ybufOff);
            for (uint32_t i = et; i < P.nybuf * stagingBytes / 4u; i += kEpiWarps * 32) yz[i] = 0u;
            fence_proxy_async_smem();
            named_bar_sync(1, kEpiWarps * 32);
            if (et == 0 && m_first < m_count) {
                int o0, o1, o2, o3;
                tile_origin(m_first, o0, o1, o2, o3);
                load_y_tile(E, E.ybuf_u32, barY, width, n_origin, o0, o1, o2, o3);
            }
        }
        E.next_origin = n_origin;
        int local = 0;
        for (int it = m_first; it < m_count; it += m_stride, ++local) {
            int o0, o1, o2, o3;
            const int mt = tile_origin(it, o0, o1, o2, o3);
            if (P.nybuf == 2) {   // y buffers and their barriers alternate with the tiles
                const uint32_t yb = local & 1u;
                E.ybuf_u32 = base + ybufOff + yb * stagingBytes, E.bar_y = barY + 8u * yb, E.y_phase = (local >> 1) & 1u;
                E.ynext_u32 = base + ybufOff + (yb ^ 1u) * stagingBytes, E.bar_ynext = barY + 8u * (yb ^ 1u);
            } else {
                E.y_phase = local & 1u;
            }
            E.has_next = it + m_stride < m_count;
            if (P.bn_y != nullptr && E.has_next) tile_origin(it + m_stride, E.n0, E.n1, E.n2, E.n3);
            const bool valid = row < rows && (o0 + i0) < P.O[0] && (o1 + i1) < P.O[1] && (o2 + i2) < P.O[2] &&
                               (o3 + i3) < P.O[3];
            const long long off = (long long)(o0 + i0) * P.os[0] + (long long)(o1 + i1) * P.os[1] +
                                  (long long)(o2 + i2) * P.os[2] + (long long)(o3 + i3) * P.os[3];
            const uint32_t buf = local & 1;
            E.bar_full = barTmemFull + 8u * buf, E.full_phase = (local >> 1) & 1u;
            const uint32_t sb = (P.nstg == 2 ? (local & 1) : 0) * stagingBytes;
            const uint32_t trow = tmem_base + buf * acc_stride + (static_cast<uint32_t>(q * 32) << 16);
            epilogue_tile(E, &mapOut, smem + stagingOff + sb, base + stagingOff + sb, statbuf, trow,
                          tmemEmptyBar + 8u * buf, width, n_origin, valid, off, o0, o1, o2, o3, mt, P.nstg == 2, q, half,
                          row, lane, et);
        }
        if (et == 0) tma_store_wait_all();
        if (P.bn_y != nullptr) {
            named_bar_sync(1, kEpiWarps * 32);   // every column owner has added its last tile
            for (int i = et; i < 2 * P.ncols; i += kEpiWarps * 32) {
                const int qn = i >= P.ncols ? 1 : 0;
                P.bn_partial[((long long)blockIdx.x * 4 + qn) * P.ncols + (i - qn * P.ncols)] = statbuf[i];
            }
        } else if (P.part_sum != nullptr) {
            named_bar_sync(1, kEpiWarps * 32);
            for (int i = et; i < 2 * P.ncols; i += kEpiWarps * 32) {
                const int qn = i >= P.ncols ? 1 : 0;
                (qn ? P.part_sq : P.part_sum)[(long long)blockIdx.x * P.part_pitch + (i - qn * P.ncols)] = statbuf[i];
            }
        }
    
      }
    }
    if (k2) cluster_sync_all();     // neither CTA leaves (or frees TMEM) while the pair still reads its shared memory
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (k2) tmem_dealloc2(tmem_base, P.tmem_cols);
        else tmem_dealloc(tmem_base, P.tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------------
// MN-major weight-gradient GEMM: ws[split][tap][ci][co] = sum_{pos in split} x[pos + tap][ci] * dy[pos][co]
// A = x panels ([pos rows][64 ci], M = 128 = two panels), B = dy panels ([pos rows][64 co]).
// grid = (panel pairs, N tiles, splits).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(192, 1)
wgrad_mnmajor_kernel(const __grid_constant__ MapPack mapsA,
                     const __grid_constant__ CUtensorMap mapB, const __grid_constant__ WgradArgs P) {
    pdl_wait();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int stages = P.stages;
    const uint32_t xpan = 2u * static_cast<uint32_t>(P.mt);            // x panels per stage
    const uint32_t stageBytes = (xpan + P.nbp) * kPanelBytes;
    const uint32_t ringBytes = stages * stageBytes;
    const uint32_t barFull = base + ringBytes;
    const uint32_t barEmpty = barFull + 8u * stages;
    const uint32_t barTmem = barEmpty + 8u * stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + ringBytes + 16u * stages + 8u);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(barFull + 8u * s, 1);
            mbar_init(barEmpty + 8u * s, 1);
        }
        mbar_init(barTmem, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), P.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int m_tile = blockIdx.x;
    const int n_tile = blockIdx.y;
    const int split = blockIdx.z;
    const int rows = P.bw * P.bh * P.bt * P.bn;
    const int kb0 = split * P.kb_per_split;
    const int kb1 = min(P.num_kb, kb0 + P.kb_per_split);
    const int p0 = 2 * P.mt * m_tile;                                  // first x panel of this CTA
    const int npan = min(static_cast<int>(xpan), P.npanels - p0);     // panels that exist (>= 1)

    // warp-uniform loops, one elected lane issues (see igemm_kmajor_kernel)
    if (warp == 0) {
        const uint32_t leader = elect_one();
        const uint32_t tx = static_cast<uint32_t>(rows) * 128u * (npan + P.nbp);
        // the (at most four) x panels of this CTA: tap geometry and channel offset are loop invariant
        Tap tapj[4];
        int c0j[4];
        const CUtensorMap* mpj[4];
        for (int j = 0; j < 4; ++j) {
            const int p = min(p0 + j, P.npanels - 1);
            const int tp = p / P.kchunks;
            c0j[j] = (p - tp * P.kchunks) << 6;
            tapj[j] = P.taps[tp];
            mpj[j] = &mapsA.m[tapj[j].map];
        }
        // position-tile coordinates advance incrementally with the k-block index
        int m = kb0;
        int iw = m % P.tw;
        m /= P.tw;
        int ih = m % P.th;
        m /= P.th;
        int it = m % P.tt;
        int in_ = m / P.tt;
        uint32_t stage = 0, phase = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(barEmpty + 8u * stage, phase ^ 1u);
            if (leader) {
                const int w0 = iw * P.bw, h0 = ih * P.bh, t0 = it * P.bt, n0 = in_ * P.bn;
                const uint32_t full = barFull + 8u * stage;
                const uint32_t sa = base + stage * stageBytes;
                mbar_expect_tx(full, tx);
                for (int j = 0; j < npan; ++j)
                    tma_load_5d(sa + j * kPanelBytes, mpj[j], full, c0j[j], w0 + tapj[j].dw, h0 + tapj[j].dh,
                                t0 + tapj[j].dt, n0);
                for (int j = 0; j < P.nbp; ++j)
                    tma_load_5d(sa + (xpan + j) * kPanelBytes, &mapB, full, n_tile * P.bn_tile + 64 * j, w0, h0, t0, n0);
            }
            __syncwarp();
            if (++stage == static_cast<uint32_t>(stages)) {
                stage = 0;
                phase ^= 1u;
            }
            if (++iw == P.tw) {
                iw = 0;
                if (++ih == P.th) {
                    ih = 0;
                    if (++it == P.tt) {
                        it = 0;
                        ++in_;
                    }
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t leader = elect_one();
        const uint32_t idesc = umma_idesc_bf16(128, P.bn_tile, 1, 1);
        const int ksteps = rows >> 4;
        // MN-major SWIZZLE_128B descriptors: LBO = panel stride (next 64 channels), SBO = 1024 B (next 8 positions)
        const uint32_t dhi = umma_desc_hi(1024, 2);
        const uint32_t lbo = (kPanelBytes >> 4) << 16;
        uint32_t stage = 0, phase = 0;
        uint32_t acc = 0;
        const int ntile = (npan + 1) >> 1;          // M tiles that have at least one panel
        for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(barFull + 8u * stage, phase);
            tc_fence_after();
            if (leader) {
                const uint32_t sa = base + stage * stageBytes;
                const uint32_t b_lo = (((sa + xpan * kPanelBytes) >> 4) & 0x3FFFu) | lbo;
                for (int t = 0; t < ntile; ++t) {   // M tiles of this CTA share the dy panels of the stage
                    const uint32_t a_lo = (((sa + 2u * t * kPanelBytes) >> 4) & 0x3FFFu) | lbo;
                    const uint32_t tacc = tmem_base + static_cast<uint32_t>(t * P.acc_stride);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {   // 16 position rows = 2048 B per step
                        if (k < ksteps) umma_bf16_lohi(tacc, a_lo + 128u * k, dhi, b_lo + 128u * k, dhi, idesc, (k == 0) ? acc : 1u);
                    }
                }
                acc = 1;
                umma_commit(barEmpty + 8u * stage);
            }
            __syncwarp();
            if (++stage == static_cast<uint32_t>(stages)) {
                stage = 0;
                phase ^= 1u;
            }
        }
        if (leader) umma_commit(barTmem);
        __syncwarp();
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        mbar_wait(barTmem, 0);
        tc_fence_after();
        for (int t = 0; t < P.mt; ++t) {
            const int p = p0 + 2 * t + (row >> 6);
            const int tp = p / P.kchunks;
            const int ci = ((p - tp * P.kchunks) << 6) + (row & 63);
            const bool valid = p < P.npanels && ci < P.ci_store;
            float* dst = P.ws + (((long long)split * P.ntaps + tp) * P.ci_pitch + ci) * P.co_pitch + n_tile * P.bn_tile;
            const uint32_t trow = tmem_base + static_cast<uint32_t>(t * P.acc_stride) + (static_cast<uint32_t>(q * 32) << 16);
            if (p0 + 2 * t >= P.npanels) break;   // warp-uniform: this M tile has no panel at all
            for (int c = 0; c < P.bn_tile; c += 16) {
                uint32_t v[16];
                tmem_ld16(trow + c, v);
                tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<uint4*>(dst + c + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, P.tmem_cols);
    }
}


// ------------------------------------------------------------------------------------------------
// Weight gradient of the temporal 3x1x1 stride-1 convolutions with a haloed activation tile.
//
// The generic kernel above gives every (tap, 64-channel chunk) panel pair its own CTA column, so the dy panels of a
// position block are re-read once per M tile and x once per tap: 1.05 GB of DRAM reads for 0.46 GB of operands on
// 144->64 (79 % DRAM utilisation).  Here one CTA owns ALL panels of its range of position blocks: the x tile carries a
// halo of one T slice on both sides, the three tap views are descriptor offsets of whole T slices into that one tile
// (the reduction index of an MN-major operand is the row, so a tap shift is a start-address offset), and the
// (tap, chunk pair) accumulators sit side by side in TMEM.  x and dy are read once (x: + 2/b3 halo).
// grid = position-block splits; ws[split][tap][ci][co].
// ------------------------------------------------------------------------------------------------
struct WgradHaloArgs {
    int32_t b[4];            // position box: inner dims W, H, N and the T extent (rows = b0*b1*b2*b3, multiple of 16)
    int32_t tl[4];
    FastDiv fd_tl0, fd_tl1, fd_tl2;
    int32_t nchunks, npairs; // 64-channel chunks of x, pairs of chunks (UMMA M = 128 = two chunks)
    int32_t ntaps;
    int32_t ci_pitch, co_pitch, bn_tile, nbp;
    int32_t stages, tmem_cols, acc_stride;
    int32_t num_kb, kb_per_split;
    uint32_t a_chunk_bytes, stage_bytes;
    // spatial 1x3x3 mode: blockIdx.y = filter column (W tap); the S = 3 filter rows are the halo taps.  With a single
    // 64-channel chunk the two halves of an M = 128 MMA are two TAPS (same buffer, one halo slice apart).
    int32_t ncopies;         // 1 (temporal) or kw (spatial)
    int32_t copy_org;        // W coordinate offset of copy 0 (= -pw)
    int32_t tap_stride;      // ws tap index = halo tap * tap_stride + copy
    int32_t pair_taps;       // 1: M halves are consecutive halo taps of chunk 0; 0: consecutive chunks of one tap
    int32_t nacc;            // accumulators per CTA
    int32_t total_taps;      // taps of the whole filter (pitch of the tap axis of ws)
    float* ws;
};

__global__ void __launch_bounds__(192, 1)
wgrad_halo_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapDy,
                  const __grid_constant__ WgradHaloArgs P) {
    pdl_wait();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int stages = P.stages;
    const uint32_t ringBytes = stages * P.stage_bytes;
    const uint32_t barFull = base + ringBytes;
    const uint32_t barEmpty = barFull + 8u * stages;
    const uint32_t barTmem = barEmpty + 8u * stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + ringBytes + 16u * stages + 8u);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(barFull + 8u * s, 1);
            mbar_init(barEmpty + 8u * s, 1);
        }
        mbar_init(barTmem, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), P.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int split = blockIdx.x;
    const int copy = blockIdx.y;
    const int kb0 = split * P.kb_per_split;
    const int kb1 = min(P.num_kb, kb0 + P.kb_per_split);
    const int inner = P.b[0] * P.b[1] * P.b[2];
    const int rows = inner * P.b[3];
    const int halo_rows = inner * (P.b[3] + P.ntaps - 1);
    const uint32_t dyOff = P.stage_bytes - static_cast<uint32_t>(P.nbp) * kPanelBytes;   // dy panels follow the x buffers

    if (warp == 0) {
        const uint32_t leader = elect_one();
        const uint32_t tx = static_cast<uint32_t>(halo_rows) * 128u * P.nchunks + static_cast<uint32_t>(rows) * 128u * P.nbp;
        uint32_t stage = 0, phase = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
            int m, m3;
            const int o0 = fdivmod(kb, P.fd_tl0, m) * P.b[0];
            const int o1 = fdivmod(m, P.fd_tl1, m) * P.b[1];
            const int o2 = fdivmod(m, P.fd_tl2, m3) * P.b[2];
            const int o3 = m3 * P.b[3];
            mbar_wait(barEmpty + 8u * stage, phase ^ 1u);
            if (leader) {
                const uint32_t full = barFull + 8u * stage;
                const uint32_t sa = base + stage * P.stage_bytes;
                mbar_expect_tx(full, tx);
                for (int c = 0; c < P.nchunks; ++c)
                    tma_load_5d(sa + c * P.a_chunk_bytes, &mapX, full, c << 6, o0 + P.copy_org + copy, o1, o2,
                                o3 - (P.ntaps >> 1));
                for (int j = 0; j < P.nbp; ++j)
                    tma_load_5d(sa + dyOff + j * kPanelBytes, &mapDy, full, 64 * j, o0, o1, o2, o3);
            }
            __syncwarp();
            if (++stage == static_cast<uint32_t>(stages)) {
                stage = 0;
                phase ^= 1u;
            }
        }
    } else if (warp == 1) {
        const uint32_t leader = elect_one();
        const uint32_t idesc = umma_idesc_bf16(128, P.bn_tile, 1, 1);
        const int ksteps = rows >> 4;
        const uint32_t dhi = umma_desc_hi(1024, 2);
        const uint32_t lboA = (P.a_chunk_bytes >> 4) << 16;     // second 64-channel half of M = the next chunk buffer
        const uint32_t lboB = (kPanelBytes >> 4) << 16;
        const uint32_t tap16 = (static_cast<uint32_t>(inner) * 128u) >> 4;   // one T slice of rows, in 16-byte units
        uint32_t stage = 0, phase = 0;
        uint32_t acc = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(barFull + 8u * stage, phase);
            tc_fence_after();
            if (leader) {
                const uint32_t sa = base + stage * P.stage_bytes;
                const uint32_t b_lo = (((sa + dyOff) >> 4) & 0x3FFFu) | lboB;
                if (P.pair_taps) {
                    // one chunk: accumulator a covers halo taps 2a and 2a+1 (LBO = one halo slice)
                    const uint32_t lboT = tap16 << 16;
                    for (int a = 0; a < P.nacc; ++a) {
                        const uint32_t a_lo = (((sa >> 4) + 2u * a * tap16) & 0x3FFFu) | lboT;
                        const uint32_t tacc = tmem_base + static_cast<uint32_t>(a * P.acc_stride);
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            if (k < ksteps)
                                umma_bf16_lohi(tacc, a_lo + 128u * k, dhi, b_lo + 128u * k, dhi, idesc, (k == 0) ? acc : 1u);
                        }
                    }
                } else {
                    int a = 0;
                    for (int t = 0; t < P.ntaps; ++t) {
                        for (int pr = 0; pr < P.npairs; ++pr, ++a) {
                            const uint32_t a_lo = ((((sa + 2u * pr * P.a_chunk_bytes) >> 4) + t * tap16) & 0x3FFFu) | lboA;
                            const uint32_t tacc = tmem_base + static_cast<uint32_t>(a * P.acc_stride);
#pragma unroll
                            for (int k = 0; k < 8; ++k) {   // 16 position rows = 2048 B per step
                                if (k < ksteps)
                                    umma_bf16_lohi(tacc, a_lo + 128u * k, dhi, b_lo + 128u * k, dhi, idesc, (k == 0) ? acc : 1u);
                            }
                        }
                    }
                }
                acc = 1;
                umma_commit(barEmpty + 8u * stage);
            }
            __syncwarp();
            if (++stage == static_cast<uint32_t>(stages)) {
                stage = 0;
                phase ^= 1u;
            }
        }
        if (leader) umma_commit(barTmem);
        __syncwarp();
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        mbar_wait(barTmem, 0);
        tc_fence_after();
        for (int a = 0; a < P.nacc; ++a) {
            int t, chunk;
            if (P.pair_taps) {
                t = 2 * a + (row >> 6);
                chunk = 0;
            } else {
                t = a / P.npairs;
                chunk = 2 * (a - t * P.npairs) + (row >> 6);
            }
            const int ci = (chunk << 6) + (row & 63);
            const bool valid = t < P.ntaps && chunk < P.nchunks && ci < P.ci_pitch;
            const int wtap = t * P.tap_stride + copy;
            float* dst = P.ws + (((long long)split * P.total_taps + wtap) * P.ci_pitch + ci) * P.co_pitch;
            const uint32_t trow = tmem_base + static_cast<uint32_t>(a * P.acc_stride) + (static_cast<uint32_t>(q * 32) << 16);
            for (int c = 0; c < P.bn_tile; c += 16) {
                uint32_t v[16];
                tmem_ld16(trow + c, v);
                tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<uint4*>(dst + c + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, P.tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------------
// Weight gradient of the wide layers (layers 2-4) on CTA pairs: cluster of 2, tcgen05 cta_group::2, M = 256.
//
// The generic kernel above computes 128 x <=256 output tiles: (128 + 256) * 2 bytes of operands per reduction row for
// 32 K MACs, i.e. 96 B/clk/SM at tensor rate against ~42 B/clk/SM the L2->SM path delivers, and on layer 4 (2 156
// positions, 72 x-panels x 1152 output channels) it needs 180 CTAs with a two-stage ring: 85-98 us against a 16 us
// tensor floor.  Here a pair of CTAs owns a 256 x <=256 tile: each CTA loads the two 64-channel x panels of ITS 128
// accumulator rows and HALF of the dy columns (the MMA reads the other half from the peer's shared memory), so
// operand bytes per SM and MAC halve; the position box is free of the 128-row limit of an M tile (rows only have to
// be padded to the K = 16 granule: 7 x 1 x 2 x 8 = 112 positions of a layer-4 tensor fill a block completely) and the
// tail rows of a padded slot are zeroed once.  Split-K only where the tiles would leave SMs idle.
// grid = (2 * M tiles, N tiles, splits), cluster (2,1,1); ws[split][tap][ci][co] as for the generic kernel.
// ------------------------------------------------------------------------------------------------
struct WgradPairArgs {
    int32_t bw, bh, bt, bn;  // box of positions forming one K block
    int32_t tw, th, tt, tn;
    int32_t rows, slot_rows; // rows of one box, rows of its shared-memory slot (multiple of 16, tail rows stay zero)
    int32_t kchunks, npanels, ntaps;
    int32_t ci_store, ci_pitch, co_pitch;
    int32_t bn_tile, half_n, nbh;   // UMMA N, columns per CTA, 64-wide dy panels per CTA
    int32_t stages, tmem_cols;
    int32_t num_kb, kb_per_split;
    float* ws;
    Tap taps[kMaxTaps];
};

__global__ void __launch_bounds__(192, 1)
wgrad_pair_kernel(const __grid_constant__ MapPack mapsA, const __grid_constant__ CUtensorMap mapB,
                  const __grid_constant__ WgradPairArgs P) {
    pdl_wait();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int stages = P.stages;
    const uint32_t rank = cluster_ctarank();
    const uint32_t panelBytes = static_cast<uint32_t>(P.slot_rows) * 128u;
    const uint32_t stageBytes = (2u + P.nbh) * panelBytes;
    const uint32_t ringBytes = stages * stageBytes;
    const uint32_t barFull = base + ringBytes;
    const uint32_t barEmpty = barFull + 8u * stages;
    const uint32_t barTmem = barEmpty + 8u * stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + ringBytes + 16u * stages + 8u);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(barFull + 8u * s, 1);
            mbar_init(barEmpty + 8u * s, 1);
        }
        mbar_init(barTmem, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc2(smem_u32(tmem_slot), P.tmem_cols);
    if (P.rows < P.slot_rows) {
        // rows of a slot past the box are never written by TMA: they are reduction rows and must read as zeros
        const uint32_t tail16 = static_cast<uint32_t>(P.slot_rows - P.rows) * 8u;   // 16-byte vectors per panel
        const uint32_t npan = stages * (2u + P.nbh);
        for (uint32_t i = threadIdx.x; i < npan * tail16; i += blockDim.x) {
            const uint32_t pn = i / tail16, v = i - pn * tail16;
            *reinterpret_cast<uint4*>(smem + pn * panelBytes + static_cast<uint32_t>(P.rows) * 128u + v * 16u) =
                make_uint4(0u, 0u, 0u, 0u);
        }
        fence_proxy_async_smem();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int m_tile = blockIdx.x >> 1;
    const int n_tile = blockIdx.y;
    const int split = blockIdx.z;
    const int kb0 = split * P.kb_per_split;
    const int kb1 = min(P.num_kb, kb0 + P.kb_per_split);
    const int p0 = 4 * m_tile + 2 * static_cast<int>(rank);     // first x panel of THIS CTA (its 128 accumulator rows)

    if (warp == 0) {
        const uint32_t leader = elect_one();
        const uint32_t tx = 2u * static_cast<uint32_t>(P.rows) * 128u * (2u + P.nbh);   // both CTAs' bytes
        const uint32_t fullLeader = mapa_shared(barFull, 0);
        // the two x panels of this CTA (a panel past the end repeats the last one; its rows are never stored)
        Tap tapj[2];
        int c0j[2];
        const CUtensorMap* mpj[2];
        for (int j = 0; j < 2; ++j) {
            const int p = min(p0 + j, P.npanels - 1);
            const int tp = p / P.kchunks;
            c0j[j] = (p - tp * P.kchunks) << 6;
            tapj[j] = P.taps[tp];
            mpj[j] = &mapsA.m[tapj[j].map];
        }
        const int nb0 = n_tile * P.bn_tile + static_cast<int>(rank) * P.half_n;   // first dy channel of this CTA
        int m = kb0;
        int iw = m % P.tw;
        m /= P.tw;
        int ih = m % P.th;
        m /= P.th;
        int it = m % P.tt;
        int in_ = m / P.tt;
        uint32_t stage = 0, phase = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(barEmpty + 8u * stage, phase ^ 1u);
            if (leader) {
                const int w0 = iw * P.bw, h0 = ih * P.bh, t0 = it * P.bt, n0 = in_ * P.bn;
                const uint32_t full = fullLeader + 8u * stage;
                const uint32_t sa = base + stage * stageBytes;
                if (rank == 0) mbar_expect_tx(barFull + 8u * stage, tx);
                for (int j = 0; j < 2; ++j)
                    tma2_load_5d(sa + j * panelBytes, mpj[j], full, c0j[j], w0 + tapj[j].dw, h0 + tapj[j].dh,
                                 t0 + tapj[j].dt, n0);
                for (int j = 0; j < P.nbh; ++j)
                    tma2_load_5d(sa + (2u + j) * panelBytes, &mapB, full, nb0 + 64 * j, w0, h0, t0, n0);
            }
            __syncwarp();
            if (++stage == static_cast<uint32_t>(stages)) {
                stage = 0;
                phase ^= 1u;
            }
            if (++iw == P.tw) {
                iw = 0;
                if (++ih == P.th) {
                    ih = 0;
                    if (++it == P.tt) {
                        it = 0;
                        ++in_;
                    }
                }
            }
        }
    } else if (warp == 1 && rank == 0) {
        const uint32_t leader = elect_one();
        const uint32_t idesc = umma_idesc_bf16(256, P.bn_tile, 1, 1);
        const int ksteps = P.slot_rows >> 4;
        // MN-major SWIZZLE_128B descriptors: LBO = panel stride (next 64 channels), SBO = 1024 B (next 8 positions)
        const uint32_t dhi = umma_desc_hi(1024, 2);
        const uint32_t lbo = (panelBytes >> 4) << 16;
        uint32_t stage = 0, phase = 0;
        uint32_t acc = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(barFull + 8u * stage, phase);
            tc_fence_after();
            if (leader) {
                const uint32_t sa = base + stage * stageBytes;
                const uint32_t a_lo = ((sa >> 4) & 0x3FFFu) | lbo;
                const uint32_t b_lo = (((sa + 2u * panelBytes) >> 4) & 0x3FFFu) | lbo;
                umma2_bf16_lohi(tmem_base, a_lo, dhi, b_lo, dhi, idesc, acc);
                for (int k = 1; k < ksteps; ++k)   // 16 position rows = 2048 B per step
                    umma2_bf16_lohi(tmem_base, a_lo + 128u * k, dhi, b_lo + 128u * k, dhi, idesc, 1u);
                acc = 1;
                umma2_commit_mc(barEmpty + 8u * stage, 3);
            }
            __syncwarp();
            if (++stage == static_cast<uint32_t>(stages)) {
                stage = 0;
                phase ^= 1u;
            }
        }
        if (leader) umma2_commit_mc(barTmem, 3);
        __syncwarp();
    } else if (warp >= 2) {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int p = p0 + (row >> 6);
        const int tp = min(p, P.npanels - 1) / P.kchunks;
        const int ci = ((p - tp * P.kchunks) << 6) + (row & 63);
        const bool valid = p < P.npanels && ci < P.ci_store;
        float* dst = P.ws + (((long long)split * P.ntaps + tp) * P.ci_pitch + ci) * P.co_pitch + n_tile * P.bn_tile;
        const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        mbar_wait(barTmem, 0);
        tc_fence_after();
        for (int c = 0; c < P.bn_tile; c += 32) {
            uint32_t v0[16], v1[16];
            const bool two = c + 16 < P.bn_tile;
            tmem_ld16(trow + c, v0);
            if (two) tmem_ld16(trow + c + 16, v1);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<uint4*>(dst + c + 4 * j) = make_uint4(v0[4 * j], v0[4 * j + 1], v0[4 * j + 2], v0[4 * j + 3]);
                if (two) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<uint4*>(dst + c + 16 + 4 * j) =
                            make_uint4(v1[4 * j], v1[4 * j + 1], v1[4 * j + 2], v1[4 * j + 3]);
                }
            }
        }
        tc_fence_before();
    }
    cluster_sync_all();     // neither CTA leaves (or frees TMEM) while the pair still reads its shared memory
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, P.tmem_cols);
    }
}

// dw[co][ci][tap] = sum_split ws[split][tap][ci][co]
// Transposing reduction through shared memory: a block owns 32 output channels x ci_tile input channels x all taps.
// The split partials are read coalesced along co, four (tap, ci) rows per warp in flight; the tile is staged in OUTPUT
// order (row = co, column j = ci * ntaps + tap, odd pitch: conflict-free both ways) and the state-dict layout is
// written in contiguous runs of ci_tile * ntaps floats per output channel.  The element-wise kernel below writes one
// 4-byte element per 32-byte sector instead (20-45 us per layer-3/4 tensor; this kernel: the time of reading the
// partials once).
__global__ void __launch_bounds__(256)
wgrad_finalize_kernel(const float* __restrict__ ws, float* __restrict__ dw, int splits, int ntaps, int ci_pitch,
                      int co_pitch, int Cin, int Cout, int ci_tile) {
    pdl_wait();
    extern __shared__ float tile[];   // [32][run | 1]
    const int run = ci_tile * ntaps;
    const int pitch = run | 1;
    const int co0 = blockIdx.x * 32;
    const int ci0 = blockIdx.y * ci_tile;
    const long long split_stride = (long long)ntaps * ci_pitch * co_pitch;
    const int lane_co = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const bool co_ok = co0 + lane_co < co_pitch;
    constexpr int kU = 8;                               // (tap, ci) rows of a warp in flight
    for (int r0 = warp; r0 < run; r0 += 8 * kU) {       // rows r0, r0+8, ... of this warp
        float acc[kU];
        const float* p[kU];
        bool ok[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int r = r0 + 8 * u;
            const int tap = r / ci_tile, ci = r - tap * ci_tile;
            acc[u] = 0.f;
            ok[u] = r < run && co_ok && ci0 + ci < ci_pitch;
            p[u] = ws + ((long long)tap * ci_pitch + ci0 + ci) * co_pitch + co0 + lane_co;
        }
        for (int sp = 0; sp < splits; ++sp) {
#pragma unroll
            for (int u = 0; u < kU; ++u)
                if (ok[u]) acc[u] += p[u][sp * split_stride];
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int r = r0 + 8 * u;
            if (r < run) {
                const int tap = r / ci_tile, ci = r - tap * ci_tile;
                tile[lane_co * pitch + ci * ntaps + tap] = acc[u];
            }
        }
    }
    __syncthreads();
    const int ci_len = min(ci_tile, Cin - ci0);
    const int len = ci_len * ntaps;   // contiguous floats per output channel
    for (int i = threadIdx.x; i < 32 * len; i += 256) {
        const int co = i / len;
        const int j = i - co * len;
        if (co0 + co < Cout) dw[((long long)(co0 + co) * Cin + ci0) * ntaps + j] = tile[co * pitch + j];
    }
}

// Element-wise variant for small weight tensors reduced over MANY splits (layer 1, stem: one thread per element
// keeps the whole GPU busy where the tiled kernel would only have a handful of blocks) and for the first-layer
// W-folded layout (tap = (dt,dh), ci = dw*8 + c).
__global__ void wgrad_finalize_small_kernel(const float* __restrict__ ws, float* __restrict__ dw, int splits, int ntaps,
                                            int ci_pitch, int co_pitch, int Cin, int Cout, int wfold_kw,
                                            int wfold_taps) {
    pdl_wait();
    const long long total = (long long)ntaps * ci_pitch * co_pitch;
    const long long split_stride = total;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int co = static_cast<int>(i % co_pitch);
        long long r = i / co_pitch;
        const int ci = static_cast<int>(r % ci_pitch);
        const int tap = static_cast<int>(r / ci_pitch);
        if (co >= Cout) continue;
        int c_true, tap_true, ntaps_true;
        if (wfold_kw > 0) {
            const int dwi = ci >> 3;
            c_true = ci & 7;
            if (dwi >= wfold_kw || c_true >= Cin) continue;
            tap_true = tap * wfold_kw + dwi;
            ntaps_true = wfold_taps;
        } else {
            if (ci >= Cin) continue;
            c_true = ci;
            tap_true = tap;
            ntaps_true = ntaps;
        }
        float acc = 0.f;
        for (int s = 0; s < splits; ++s) acc += ws[s * split_stride + i];
        dw[((long long)co * Cin + c_true) * ntaps_true + tap_true] = acc;
    }
}

// weight packing: fp32 [Cout][Cin][taps] -> bf16 fprop image [tap][Cout][kpitch] and dgrad image [tap][Cin][cpad(Cout)]
__global__ void pack_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf,
                                   __nv_bfloat16* __restrict__ wd, int Cout, int Cin, int ntaps, int kpitch,
                                   int copitch, int wfold_kw) {
    pdl_wait();
    const int ftaps = wfold_kw > 0 ? ntaps / wfold_kw : ntaps;
    const long long nf = wf ? (long long)ftaps * Cout * kpitch : 0;
    const long long nd = wd ? (long long)ntaps * Cin * copitch : 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nf + nd;
         i += (long long)gridDim.x * blockDim.x) {
        if (i < nf) {
            const int k = static_cast<int>(i % kpitch);
            long long r = i / kpitch;
            const int co = static_cast<int>(r % Cout);
            const int tap = static_cast<int>(r / Cout);
            float v = 0.f;
            if (wfold_kw > 0) {
                const int dwi = k >> 3, c = k & 7;
                if (dwi < wfold_kw && c < Cin) v = w[((long long)co * Cin + c) * ntaps + tap * wfold_kw + dwi];
            } else if (k < Cin) {
                v = w[((long long)co * Cin + k) * ntaps + tap];
            }
            wf[i] = __float2bfloat16(v);
        } else {
            const long long j = i - nf;
            const int co = static_cast<int>(j % copitch);
            long long r = j / copitch;
            const int ci = static_cast<int>(r % Cin);
            const int tap = static_cast<int>(r / Cin);
            const float v = co < Cout ? w[((long long)co * Cin + ci) * ntaps + tap] : 0.f;
            wd[j] = __float2bfloat16(v);
        }
    }
}

// All convolutions of a network in ONE launch (the per-layer kernels are 5-15 us of mostly launch latency each).
constexpr int kMaxPackItems = 32;   // PackBatch travels as a kernel parameter (< 4 KB)
struct PackItem {
    const float* w;
    // optional inference-time BatchNorm folding (running statistics): w'[co] = w[co] * gamma/sqrt(var+eps),
    // bias_out[co] = beta - mean * gamma/sqrt(var+eps)
    const float* gamma;
    const float* beta;
    const float* rmean;
    const float* rvar;
    float* bias_out;
    float eps;
    __nv_bfloat16* wf;
    __nv_bfloat16* wd;
    int32_t Cout, Cin, ntaps, kpitch, copitch, wfold_kw;
    long long start;   // first flat element index of this item; [start, start + nf) fprop image, then the dgrad image,
    long long nf;      // then (folding) Cout bias entries
    long long nd;
};
struct PackBatch {
    int32_t n;
    long long total;
    PackItem it[kMaxPackItems];
};

__global__ void __launch_bounds__(256)
pack_weights_batched_kernel(const __grid_constant__ PackBatch B) {
    pdl_wait();
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < B.total;
         g += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = B.n - 1;   // last item with start <= g
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (B.it[mid].start <= g) lo = mid;
            else hi = mid - 1;
        }
        const PackItem& I = B.it[lo];
        const long long i = g - I.start;
        const bool fold = I.rvar != nullptr;
        if (i < I.nf) {
            const int k = static_cast<int>(i % I.kpitch);
            long long r = i / I.kpitch;
            const int co = static_cast<int>(r % I.Cout);
            const int tap = static_cast<int>(r / I.Cout);
            float v = 0.f;
            if (I.wfold_kw > 0) {
                const int dwi = k >> 3, c = k & 7;
                if (dwi < I.wfold_kw && c < I.Cin) v = I.w[((long long)co * I.Cin + c) * I.ntaps + tap * I.wfold_kw + dwi];
            } else if (k < I.Cin) {
                v = I.w[((long long)co * I.Cin + k) * I.ntaps + tap];
            }
            if (fold) v *= (I.gamma ? I.gamma[co] : 1.f) * rsqrtf(I.rvar[co] + I.eps);
            I.wf[i] = __float2bfloat16(v);
        } else if (i < I.nf + I.nd) {
            const long long j = i - I.nf;
            const int co = static_cast<int>(j % I.copitch);
            long long r = j / I.copitch;
            const int ci = static_cast<int>(r % I.Cin);
            const int tap = static_cast<int>(r / I.Cin);
            const float v = co < I.Cout ? I.w[((long long)co * I.Cin + ci) * I.ntaps + tap] : 0.f;
            I.wd[j] = __float2bfloat16(v);
        } else {
            const int co = static_cast<int>(i - I.nf - I.nd);
            const float sc = (I.gamma ? I.gamma[co] : 1.f) * rsqrtf(I.rvar[co] + I.eps);
            I.bias_out[co] = (I.beta ? I.beta[co] : 0.f) - I.rmean[co] * sc;
        }
    }
}

// Tiled variant for plain (not W-folded) weights: one block stages 16 output channels x one chunk of input channels x
// all taps -- 16 contiguous runs of the fp32 master weight, read coalesced -- as bf16 in shared memory and writes both
// images with 16-byte stores (8 consecutive k of the fprop image, 8 consecutive output channels of the dgrad image).
// The element-wise kernel above spends two 64-bit divisions per element and reads the dgrad image's sources one
// 32-byte sector per element: 375 us per optimizer step for R(2+1)D-18, against ~50 us of HBM time for the 250 MB moved.
constexpr int kPackCo = 16;
constexpr int kPackElems = 1024;   // (input channel, tap) elements staged per output channel
struct PackTileItem {
    const float* w;
    const float* gamma;
    const float* beta;
    const float* rmean;
    const float* rvar;
    float* bias_out;
    __nv_bfloat16* wf;
    __nv_bfloat16* wd;
    float eps;
    int32_t Cout, Cin, ntaps, kpitch, copitch;
    int32_t ci_chunk, n_ci;   // input channels per block, blocks along the input channels
    int32_t unit_start;       // first block of this item
};
struct PackTileBatch {
    int32_t n, units;
    PackTileItem it[kMaxPackItems];
};

__global__ void __launch_bounds__(256)
pack_weights_tiled_kernel(const __grid_constant__ PackTileBatch B) {
    pdl_wait();
    __shared__ __align__(16) __nv_bfloat16 tile[kPackCo][kPackElems + 8];
    int lo = 0, hi = B.n - 1;   // last item with unit_start <= blockIdx.x
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (B.it[mid].unit_start <= static_cast<int>(blockIdx.x)) lo = mid;
        else hi = mid - 1;
    }
    const PackTileItem& I = B.it[lo];
    const int unit = static_cast<int>(blockIdx.x) - I.unit_start;
    const int co0 = (unit / I.n_ci) * kPackCo;
    const int cc = unit % I.n_ci;
    const int ci0 = cc * I.ci_chunk;
    const int ci_len = min(I.ci_chunk, I.Cin - ci0);
    const int len = ci_len * I.ntaps;
    const int tid = threadIdx.x;
    const bool fold = I.rvar != nullptr;
    // stage: rows = output channels, columns = (ci - ci0) * ntaps + tap, exactly the master weight's order
    for (int r = 0; r < kPackCo; ++r) {
        const int co = co0 + r;
        if (co < I.Cout) {
            const float sc = fold ? (I.gamma ? I.gamma[co] : 1.f) * rsqrtf(I.rvar[co] + I.eps) : 1.f;
            const float* src = I.w + ((long long)co * I.Cin + ci0) * I.ntaps;
            for (int e = tid; e < len; e += 256) tile[r][e] = __float2bfloat16(fold ? src[e] * sc : src[e]);
        } else {
            for (int e = tid; e < len; e += 256) tile[r][e] = __float2bfloat16(0.f);
        }
    }
    if (fold && cc == 0 && tid < kPackCo && co0 + tid < I.Cout) {
        const int co = co0 + tid;
        const float sc = (I.gamma ? I.gamma[co] : 1.f) * rsqrtf(I.rvar[co] + I.eps);
        I.bias_out[co] = (I.beta ? I.beta[co] : 0.f) - I.rmean[co] * sc;
    }
    __syncthreads();
    const uint16_t* t16 = reinterpret_cast<const uint16_t*>(&tile[0][0]);
    constexpr int kRow = kPackElems + 8;
    if (I.wf != nullptr) {
        // fprop image [tap][Cout][kpitch]: this block's k range, zero beyond Cin (the last chunk also writes the pad lanes)
        const int k_end = (cc + 1 == I.n_ci) ? I.kpitch : ci0 + I.ci_chunk;
        const int kv = (k_end - ci0) >> 3;                  // 16-byte vectors per (tap, co)
        const int rows = min(kPackCo, I.Cout - co0);
        const int total = I.ntaps * rows * kv;
        for (int idx = tid; idx < total; idx += 256) {
            const int v = idx % kv;
            const int rr = (idx / kv) % rows;
            const int tap = idx / (kv * rows);
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k0 = 8 * v + 2 * j;               // relative to ci0
                const uint32_t lo16 = k0 < ci_len ? t16[rr * kRow + k0 * I.ntaps + tap] : 0u;
                const uint32_t hi16 = k0 + 1 < ci_len ? t16[rr * kRow + (k0 + 1) * I.ntaps + tap] : 0u;
                pk[j] = lo16 | (hi16 << 16);
            }
            __nv_bfloat16* dst = I.wf + ((long long)tap * I.Cout + co0 + rr) * I.kpitch + ci0 + 8 * v;
            *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
    }
    if (I.wd != nullptr) {
        // dgrad image [tap][Cin][copitch]: 8 consecutive output channels per store (rows past Cout were staged as zeros)
        const int halves = min(2, (I.copitch - co0) >> 3);
        const int total = I.ntaps * ci_len * halves;
        for (int idx = tid; idx < total; idx += 256) {
            const int h = idx % halves;
            const int ci = (idx / halves) % ci_len;
            const int tap = idx / (halves * ci_len);
            const int col = ci * I.ntaps + tap;
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                pk[j] = static_cast<uint32_t>(t16[(8 * h + 2 * j) * kRow + col]) |
                        (static_cast<uint32_t>(t16[(8 * h + 2 * j + 1) * kRow + col]) << 16);
            __nv_bfloat16* dst = I.wd + ((long long)tap * I.Cin + ci0 + ci) * I.copitch + co0 + 8 * h;
            *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host-side geometry
// ------------------------------------------------------------------------------------------------
struct Shape {
    int To, Ho, Wo;
    int cinp, coutp;
    bool wfold;
    int keff;    // reduction channels per tap on the x side (Cin; wfold: the W window, 8 * kw rounded up to 16)
    int kpitch;  // channel pitch of the fprop weight image
    int ntaps;   // kt*kh*kw
    int ftaps;   // taps seen by fprop/wgrad kernels (kw folded away for wfold)
};

int check_desc(const zsv_conv_desc* d, Shape* s) {
    if (!d) return fail(ZSV_ERR_BAD_ARG, "null conv descriptor");
    if (d->N < 1 || d->T < 1 || d->H < 1 || d->W < 1 || d->Cin < 1 || d->Cout < 1)
        return fail(ZSV_ERR_BAD_ARG, "conv desc: non-positive extent");
    if (d->kt < 1 || d->kh < 1 || d->kw < 1 || d->kt > 7 || d->kh > 7 || d->kw > 7)
        return fail(ZSV_ERR_BAD_ARG, "conv desc: filter extents must be in 1..7");
    const int st[3] = {d->st, d->sh, d->sw};
    for (int i = 0; i < 3; ++i)
        if (st[i] != 1 && st[i] != 2) return fail(ZSV_ERR_UNSUPPORTED, "conv desc: stride must be 1 or 2");
    if (d->pt < 0 || d->ph < 0 || d->pw < 0) return fail(ZSV_ERR_BAD_ARG, "conv desc: negative padding");
    s->To = (d->T + 2 * d->pt - d->kt) / d->st + 1;
    s->Ho = (d->H + 2 * d->ph - d->kh) / d->sh + 1;
    s->Wo = (d->W + 2 * d->pw - d->kw) / d->sw + 1;
    if (s->To < 1 || s->Ho < 1 || s->Wo < 1) return fail(ZSV_ERR_BAD_ARG, "conv desc: empty output");
    s->cinp = cpad(d->Cin);
    s->coutp = cpad(d->Cout);
    s->ntaps = d->kt * d->kh * d->kw;
    s->wfold = d->x_layout == ZSV_CONV_X_WFOLD;
    if (s->wfold) {
        if (s->cinp != 8 || d->kw > 8) return fail(ZSV_ERR_UNSUPPORTED, "wfold layout needs Cin <= 8 and kw <= 8");
        if ((s->Wo - 1) * d->sw + 8 > d->W + kWfoldWpad || d->pw > kWfoldWpad)
            return fail(ZSV_ERR_UNSUPPORTED, "wfold layout: window exceeds padded row");
        // the W window of a tap is kw positions x 8 channel lanes of the 64-element row: k-steps (and TMA bytes) past it
        // would multiply zeros (C3D's 3x3x3 first layer: K = 32 per tap instead of 64)
        s->keff = std::min(64, (8 * d->kw + 15) & ~15);
        s->kpitch = 64;
        s->ftaps = d->kt * d->kh;
    } else if (d->x_layout == ZSV_CONV_X_NDHWC) {
        s->keff = d->Cin;
        s->kpitch = s->cinp;
        s->ftaps = s->ntaps;
    } else {
        return fail(ZSV_ERR_BAD_ARG, "conv desc: unknown x_layout %d", d->x_layout);
    }
    if (s->ftaps > kMaxTaps) return fail(ZSV_ERR_UNSUPPORTED, "conv desc: more than %d taps", kMaxTaps);
    return ZSV_OK;
}

struct Box {
    int bw, bh, bt, bn;
    int rows() const { return bw * bh * bt * bn; }
};

// Pick the box of positions (<= 128 rows) covering (OW,OH,OT,ON) with the least wasted MMA rows.
// rows16: rows must be a multiple of 16 (wgrad consumes rows as the K dimension).
Box choose_box(int OW, int OH, int OT, int ON, bool rows16) {
    if (const char* e = getenv("ZSV_DEBUG_BOX")) {   // tuning aid: force the position box "bw,bh,bt,bn"
        Box b{1, 1, 1, 1};
        if (sscanf(e, "%d,%d,%d,%d", &b.bw, &b.bh, &b.bt, &b.bn) == 4 && b.rows() <= 128 && (!rows16 || b.rows() % 16 == 0))
            return b;
    }
    Box best{1, 1, 1, 1};
    double best_score = -1.0;
    const double total = (double)OW * OH * OT * ON;
    for (int bw = 1; bw <= std::min(OW, 128); ++bw)
        for (int bh = 1; bh <= std::min(OH, 128 / bw); ++bh)
            for (int bt = 1; bt <= std::min(OT, 128 / (bw * bh)); ++bt)
                for (int bn = 1; bn <= std::min(ON, 128 / (bw * bh * bt)); ++bn) {
                    const int rows = bw * bh * bt * bn;
                    if (rows16 && (rows & 15)) continue;
                    const double tiles = (double)ceil_div(OW, bw) * ceil_div(OH, bh) * ceil_div(OT, bt) *
                                         ceil_div(ON, bn);
                    const double cost = rows16 ? tiles * (rows + 16) : tiles * 128.0;
                    // tiny bias towards wide-in-w, then tall boxes (longer contiguous runs in memory)
                    const double score = total / cost + 1e-6 * bw + 1e-8 * bh;
                    if (score > best_score) {
                        best_score = score;
                        best = Box{bw, bh, bt, bn};
                    }
                }
    if (rows16 && best_score < 0) {
        // tensor smaller than 16 positions in every factorisation: over-cover with a padded box (OOB rows read zero)
        best = Box{std::min(128, 16 * ceil_div(OW, 16)), 1, 1, 1};
        if (best.bw > 128) best.bw = 128;
    }
    return best;
}

// N tile (multiple of 16, <= 256) for `cols` output channels given the number of M tiles.
void choose_ntile(int cols, long long m_tiles, int* bn_tile, int* n_tiles, int max_bn = 256) {
    const int cols16 = (cols + 15) & ~15;
    const int sms = std::max(1, sm_count());
    double best = 1e300;
    int best_bn = std::min(cols16, max_bn), best_n = ceil_div(cols16, std::min(cols16, max_bn));
    for (int nt = ceil_div(cols16, max_bn); nt <= ceil_div(cols16, max_bn) + 6 && nt <= ceil_div(cols16, 16); ++nt) {
        int bn = ((ceil_div(cols16, nt) + 15) & ~15);
        if (nt > 1) bn = (bn + 63) & ~63;   // the output is stored in 64-channel panels: tile origins stay panel-aligned
        if (bn > max_bn) continue;
        if (nt > 1 && (long long)bn * (nt - 1) >= cols16) continue;   // last tile would be empty
        if (bn < 64 && cols16 >= 64) continue;
        const long long ctas = m_tiles * nt;
        const long long slots = sms;
        const long long waves = ceil_div_ll(ctas, slots);
        // per-CTA cost ~ A-tile handling (fixed) + MMA/B/epilogue work proportional to bn
        const double cost = (double)waves * (bn + 48.0);
        if (cost < best) {
            best = cost;
            best_bn = bn;
            best_n = nt;
        }
    }
    *bn_tile = best_bn;
    *n_tiles = best_n;
}

int pow2_cols(int n) {
    int c = 32;
    while (c < n) c <<= 1;
    return c;
}

struct DimTap {
    int j;       // filter index along this dimension
    int parity;  // parity plane of the source coordinate (0 when stride 1)
    int off;     // coordinate offset in the (plane) index space
};

// forward taps along one dimension: source i = s*o + j - p
std::vector<DimTap> fwd_dim_taps(int k, int s, int p) {
    std::vector<DimTap> v;
    for (int j = 0; j < k; ++j) {
        const int e = j - p;
        if (s == 1) {
            v.push_back({j, 0, e});
        } else {
            const int par = ((e % 2) + 2) % 2;
            v.push_back({j, par, (e - par) / 2});
        }
    }
    return v;
}
// dgrad taps along one dimension for dx positions i = s*q + cls: source o = (i + p - j)/s when divisible
std::vector<DimTap> bwd_dim_taps(int k, int s, int p, int cls) {
    std::vector<DimTap> v;
    for (int j = 0; j < k; ++j) {
        const int e = cls + p - j;
        if (s == 1) {
            v.push_back({j, 0, e});
        } else if (((e % 2) + 2) % 2 == 0) {
            v.push_back({j, 0, e / 2});
        }
    }
    return v;
}

// Activation tensor map for parity plane (pt_, ph_, pw_) of x with box (64, bw, bh, bt, bn).
int make_x_map(CUtensorMap* m, const zsv_conv_desc* d, const Shape& s, const void* x, int pt_, int ph_, int pw_,
               const Box& b) {
    uint64_t dims[5], strides[4];
    uint32_t box[5] = {64, (uint32_t)b.bw, (uint32_t)b.bh, (uint32_t)b.bt, (uint32_t)b.bn};
    const char* p = static_cast<const char*>(x);
    if (s.wfold) {
        const long long Wp = d->W + kWfoldWpad;
        const long long rowB = Wp * 8 * 2;
        // column (pw - d->pw .. ) : repack placed exactly d->pw zero columns on the left, so window start = sw*wo
        dims[0] = s.keff;
        dims[1] = s.Wo;
        dims[2] = (d->H - ph_ + d->sh - 1) / d->sh;
        dims[3] = (d->T - pt_ + d->st - 1) / d->st;
        dims[4] = d->N;
        strides[0] = (uint64_t)d->sw * 8 * 2;
        strides[1] = (uint64_t)rowB * d->sh;
        strides[2] = (uint64_t)rowB * d->H * d->st;
        strides[3] = (uint64_t)rowB * d->H * d->T;
        p += ((long long)pt_ * d->H + ph_) * rowB;
    } else {
        const long long cB = (long long)s.cinp * 2;
        dims[0] = d->Cin;
        dims[1] = (d->W - pw_ + d->sw - 1) / d->sw;
        dims[2] = (d->H - ph_ + d->sh - 1) / d->sh;
        dims[3] = (d->T - pt_ + d->st - 1) / d->st;
        dims[4] = d->N;
        strides[0] = (uint64_t)cB * d->sw;
        strides[1] = (uint64_t)cB * d->W * d->sh;
        strides[2] = (uint64_t)cB * d->W * d->H * d->st;
        strides[3] = (uint64_t)cB * d->W * d->H * d->T;
        p += (((long long)pt_ * d->H + ph_) * d->W + pw_) * cB;
    }
    for (int i = 1; i < 4; ++i)
        if (dims[i] == 0) dims[i] = 1;
    return make_map(m, p, 5, dims, strides, box);
}

// Plain NDHWC map of an activation-shaped tensor [N][T][H][W][pitch] with C valid channels.
int make_plain_map(CUtensorMap* m, const void* base, int N, int T, int H, int W, int C, int pitch, const Box& b) {
    uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)T, (uint64_t)N};
    const uint64_t cB = (uint64_t)pitch * 2;
    uint64_t strides[4] = {cB, cB * W, cB * W * H, cB * W * H * T};
    uint32_t box[5] = {64, (uint32_t)b.bw, (uint32_t)b.bh, (uint32_t)b.bt, (uint32_t)b.bn};
    return make_map(m, base, 5, dims, strides, box);
}

// Forward-style tap table (used by fprop and wgrad): fills taps + the parity maps they reference.
int build_fwd_taps(const zsv_conv_desc* d, const Shape& s, const void* x, const Box& b, Tap* taps, CUtensorMap* maps,
                   int* nmaps) {
    const int kw_eff = s.wfold ? 1 : d->kw;
    std::vector<DimTap> tt = fwd_dim_taps(d->kt, d->st, d->pt);
    std::vector<DimTap> th = fwd_dim_taps(d->kh, d->sh, d->ph);
    std::vector<DimTap> tw = s.wfold ? std::vector<DimTap>{{0, 0, 0}} : fwd_dim_taps(d->kw, d->sw, d->pw);
    int keys[kMaxMaps];
    *nmaps = 0;
    for (const DimTap& a : tt)
        for (const DimTap& bq : th)
            for (const DimTap& c : tw) {
                const int key = (a.parity << 2) | (bq.parity << 1) | c.parity;
                int id = -1;
                for (int i = 0; i < *nmaps; ++i)
                    if (keys[i] == key) id = i;
                if (id < 0) {
                    if (*nmaps == kMaxMaps) return fail(ZSV_ERR_UNSUPPORTED, "conv needs more than %d parity maps", kMaxMaps);
                    id = (*nmaps)++;
                    keys[id] = key;
                    int rc = make_x_map(&maps[id], d, s, x, a.parity, bq.parity, c.parity, b);
                    if (rc) return rc;
                }
                const int ti = (a.j * d->kh + bq.j) * kw_eff + c.j;
                taps[ti].map = (int16_t)id;
                taps[ti].dt = (int16_t)a.off;
                taps[ti].dh = (int16_t)bq.off;
                taps[ti].dw = (int16_t)c.off;
                taps[ti].btap = (int16_t)ti;
            }
    for (int i = *nmaps; i < kMaxMaps; ++i) maps[i] = maps[0];
    return ZSV_OK;
}

// host-side view of a zsv_bn_bwd_fuse request while the (one or more) launches of a dgrad call are issued
struct BnFuseLaunch {
    const __nv_bfloat16* y;   // already offset to the parity class of the launch
    const float4* tab;
    int relu;
    float* partial;
    int capacity;             // rows
    int rows_used;
};

// shared-memory scratch of the per-CTA running column sums: [2][ncols] fp32 (BatchNorm statistics in fprop,
// (sum dz, sum dz*y) in the fused BatchNorm backward)
int bn_scratch_bytes(int ncols, int bn_tile) { return (((2 * ncols + 8 * bn_tile) * 4) + 1023) & ~1023; }

// CTA pairs (cta_group::2) for the generic kernel
bool igemm_use_pair(int bn_tile, long long m_tiles) {
    // on by default since the cluster-scope "accumulator drained" arrive is relaxed (profiles/r01_halo_pair_ab.txt:
    // +1.5% on the whole training step); ZSV_2CTA=0 runs the single-CTA kernel
    const char* e = getenv("ZSV_2CTA");
    return !(e && atoi(e) == 0) && (bn_tile % 16) == 0 && m_tiles >= 2;
}
// Tiles of at most 64 output channels with two staging buffers run the split epilogue (convert / finish groups on
// consecutive tiles); ZSV_EPI_SPLIT=0 keeps the single group everywhere.
bool epilogue_split(int bn_tile, int nstg) {
    const char* e = getenv("ZSV_EPI_SPLIT");
    return !(e && atoi(e) == 0) && bn_tile <= 64 && nstg == 2;
}
// grid of the generic kernel (one BatchNorm partial row per CTA)
int igemm_grid(int bn_tile, long long m_tiles, int n_tiles) {
    if (igemm_use_pair(bn_tile, m_tiles))
        return 2 * (int)std::min<long long>(((m_tiles + 1) / 2) * n_tiles, sm_count() / 2);
    return (int)std::min<long long>(m_tiles * n_tiles, sm_count());
}

int launch_igemm(const CUtensorMap* maps, const CUtensorMap& mapB, const CUtensorMap& mapOut, IgemmArgs& a,
                 long long m_tiles, int n_tiles, cudaStream_t stream, BnFuseLaunch* fuse = nullptr,
                 const CUtensorMap* mapY = nullptr) {
    // one persistent CTA per SM owns (almost) all shared memory: as many ring stages as fit, at most 8
    // two output staging buffers when at least 3 ring stages still fit beside them
    if (fuse && !mapY) return fail(ZSV_ERR_BAD_ARG, "igemm: BN-backward fusion without a tensor map for y");
    const int scratch = (fuse || a.part_sum) ? bn_scratch_bytes(a.ncols, 0) : 0;
    // fused BatchNorm backward: one more tile-sized buffer, for y
    const int ybuf = fuse ? ((a.bn_tile + 63) / 64) * (int)kPanelBytes : 0;
    // shared memory per ring stage: the A tile + the B rows THIS CTA holds (half of them in a CTA pair); the staging
    // buffers always hold the full N tile
    const bool pair = igemm_use_pair(a.bn_tile, m_tiles);
    const int b_rows = pair ? a.bn_tile / 2 : a.bn_tile;
    int nybuf = fuse ? 1 : 0;
    auto smem_for = [&](int stages_, int nstg_) {
        return 1024 + stages_ * ((int)kPanelBytes + b_rows * 128) + nstg_ * ((a.bn_tile + 63) / 64) * (int)kPanelBytes +
               nybuf * ybuf + scratch + 16 * stages_ + 48 + 64;
    };
    int nstg = 2;
    if (const char* e = getenv("ZSV_DEBUG_NSTG")) nstg = atoi(e) == 1 ? 1 : 2;
    int stages = 8;
    while (stages > 2 && smem_for(stages, nstg) > 226 * 1024) --stages;
    // a second staging buffer only overlaps the TMA store of a tile with the next epilogue; ring depth hides load
    // latency for every k-block, so it wins when the two compete (wide N tiles)
    if (nstg == 2 && stages < 5) {
        int s1 = 8;
        while (s1 > 2 && smem_for(s1, 1) > 226 * 1024) --s1;
        if (s1 > stages && !getenv("ZSV_DEBUG_KEEP_NSTG2")) nstg = 1, stages = s1;
    }
    // fused BatchNorm backward: a second y buffer (y requested a whole tile ahead) if the ring keeps >= 4 stages and
    // loses at most one
    if (fuse && !(getenv("ZSV_NYBUF") && atoi(getenv("ZSV_NYBUF")) == 1)) {
        nybuf = 2;
        int s2 = 8;
        while (s2 > 2 && smem_for(s2, nstg) > 226 * 1024) --s2;
        if (s2 >= 4 && s2 >= stages - 1 && smem_for(s2, nstg) <= 226 * 1024) stages = s2;
        else nybuf = 1;
    }
    if (const char* e = getenv("ZSV_DEBUG_STAGES")) stages = std::max(2, std::min(stages, atoi(e)));
    a.nstg = nstg;
    a.nybuf = nybuf;
    a.stages = stages;
    a.tmem_cols = 2 * pow2_cols(a.bn_tile);   // two accumulator buffers
    if (a.tmem_cols > 512) return fail(ZSV_ERR_UNSUPPORTED, "igemm: N tile %d too wide for two TMEM buffers", a.bn_tile);
    const int smem = smem_for(stages, nstg);
    if (smem > 227 * 1024) return fail(ZSV_ERR_UNSUPPORTED, "igemm: shared memory budget exceeded (%d bytes)", smem);
    a.scratch_bytes = scratch;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(igemm_kmajor_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (attr_err == cudaSuccess)
            attr_err = cudaFuncSetAttribute(igemm_kmajor_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (attr_err == cudaSuccess)
            attr_err = cudaFuncSetAttribute(igemm_kmajor_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (attr_err == cudaSuccess)
            attr_err = cudaFuncSetAttribute(igemm_kmajor_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    if (attr_err != cudaSuccess)
        return fail(ZSV_ERR_CUDA, "cudaFuncSetAttribute(igemm) failed: %s", cudaGetErrorString(attr_err));
    if (m_tiles * n_tiles > 0x7fffffffLL) return fail(ZSV_ERR_UNSUPPORTED, "too many tiles");
    a.m_tiles = (int)m_tiles;
    a.n_tiles = n_tiles;
    a.fd_ntiles = make_fastdiv(n_tiles), a.fd_tw = make_fastdiv(a.tw), a.fd_th = make_fastdiv(a.th);
    a.fd_tt = make_fastdiv(a.tt);
    if (const char* e = getenv("ZSV_DEBUG_EPI")) a.debug = atoi(e);
    const bool two = igemm_use_pair(a.bn_tile, m_tiles);
    const long long tiles = two ? ((m_tiles + 1) / 2) * n_tiles : m_tiles * n_tiles;
    const int grid = two ? 2 * (int)std::min<long long>(tiles, sm_count() / 2) : (int)std::min<long long>(tiles, sm_count());
    if (fuse) {
        if (fuse->rows_used + grid > fuse->capacity)
            return fail(ZSV_ERR_WORKSPACE, "dgrad: BN-fusion partial buffer holds %d rows, need %d", fuse->capacity,
                        fuse->rows_used + grid);
        a.bn_y = fuse->y, a.bn_tab = fuse->tab, a.bn_relu = fuse->relu;
        a.bn_partial = fuse->partial + (size_t)fuse->rows_used * 4 * a.ncols;
        fuse->rows_used += grid;
    }
    MapPack pack;
    for (int i = 0; i < kMaxMaps; ++i) pack.m[i] = maps[i];
    const bool split = epilogue_split(a.bn_tile, a.nstg);
    if (two) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(grid), cfg.blockDim = dim3(kIgemmThreads), cfg.dynamicSmemBytes = smem, cfg.stream = stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
        pdl_attribute(&attr[1]);
        cfg.attrs = attr, cfg.numAttrs = 2;
        cudaError_t e = split ? cudaLaunchKernelEx(&cfg, igemm_kmajor_kernel<true, true>, pack, mapB, mapOut, mapY ? *mapY : mapOut, a)
                              : cudaLaunchKernelEx(&cfg, igemm_kmajor_kernel<true, false>, pack, mapB, mapOut, mapY ? *mapY : mapOut, a);
        if (e != cudaSuccess) return fail(ZSV_ERR_CUDA, "launch of igemm_kmajor_kernel<pair> failed: %s", cudaGetErrorString(e));
    } else if (split) {
        zsv::launch(igemm_kmajor_kernel<false, true>, grid, kIgemmThreads, smem, stream, pack, mapB, mapOut, mapY ? *mapY : mapOut, a);
    } else {
        zsv::launch(igemm_kmajor_kernel<false, false>, grid, kIgemmThreads, smem, stream, pack, mapB, mapOut, mapY ? *mapY : mapOut, a);
    }
    ZSV_LAUNCH_CHECK("igemm_kmajor_kernel");
    return ZSV_OK;
}

// ---- halo kernel planning -------------------------------------------------------------------------
struct HaloPlan {
    bool ok;
    bool spatial;       // shift dim = H (taps along W are copies) ; otherwise shift dim = T
    bool pair;          // CTA-pair variant: M = 256 tiles, half of the weight rows resident per CTA
    bool wshift;        // W taps by descriptor offset into one widened box (spatial, b[0] == 8)
    int b[4], tl[4], O[4];
    int S, ncopies, stages, nstg, nybuf, bn_tile, n_step, n_tiles, nchunks, tail_box;
    long long m_tiles;
    uint32_t a_stage_bytes, b_main_bytes, b_tail_bytes, b_total_bytes;
    int smem;
    int scratch;
};

inline uint32_t align1k(uint32_t v) { return (v + 1023u) & ~1023u; }

// act extents (W,H,T,N) of the GEMM-M space (stride-1 conv: output extents == input extents), reduction channels
// kdim, output columns `cols`, filter (kt,kh,kw).  Only spatial 1xkhxkw (kh==3) and temporal ktx1x1 (kt==3).
// scratch_mode: 0 = none, 1 = BatchNorm statistics (fprop), 2 = fused BatchNorm backward (dgrad)
HaloPlan plan_halo_impl(int W, int H, int T, int N, int kdim, int cols, int kt, int kh, int kw, int scratch_mode, bool pair) {
    HaloPlan p;
    memset(&p, 0, sizeof(p));
    p.pair = pair;
    if (getenv("ZSV_DEBUG_NO_HALO")) return p;
    if (kt == 1 && kh == 3 && kw >= 1 && kw <= kMaxCopies) {
        p.spatial = true;
        p.S = kh;
        p.ncopies = kw;
    } else if (kt == 3 && kh == 1 && kw == 1) {
        p.spatial = false;
        p.S = kt;
        p.ncopies = 1;
    } else {
        return p;
    }
    // box order: spatial (W, T, N | H), temporal (W, H, N | T)
    const int E[4] = {W, p.spatial ? T : H, N, p.spatial ? H : T};
    double best = -1;
    for (int b0 = 1; b0 <= std::min(E[0], 128); ++b0)
        for (int b1 = 1; b1 <= std::min(E[1], 128 / b0); ++b1)
            for (int b2 = 1; b2 <= std::min(E[2], 128 / (b0 * b1)); ++b2) {
                const int inner = b0 * b1 * b2;
                if (inner % 8) continue;
                for (int b3 = 1; b3 <= std::min(E[3], 128 / inner); ++b3) {
                    if (b3 + p.S - 1 > 256) continue;
                    const double tiles = (double)ceil_div(E[0], b0) * ceil_div(E[1], b1) * ceil_div(E[2], b2) *
                                         ceil_div(E[3], b3);
                    const double eff = ((double)E[0] * E[1] * E[2] * E[3]) / (tiles * 128.0);
                    const double halo = (double)(b3 + p.S - 1) / b3;
                    const double score = eff / (0.5 + 0.5 * halo) + 1e-6 * b0;   // MMA rows wasted vs bytes re-read
                    if (score > best) {
                        best = score;
                        p.b[0] = b0, p.b[1] = b1, p.b[2] = b2, p.b[3] = b3;
                    }
                }
            }
    if (const char* e = getenv("ZSV_DEBUG_HALO_BOX")) {   // tuning aid: force the box "b0,b1,b2,b3"
        int q[4];
        if (sscanf(e, "%d,%d,%d,%d", &q[0], &q[1], &q[2], &q[3]) == 4 &&
            (q[0] * q[1] * q[2]) % 8 == 0 && q[0] * q[1] * q[2] * q[3] <= 128)
            for (int i = 0; i < 4; ++i) p.b[i] = q[i];
        best = 1;
    }
    if (best < 0) return p;
    for (int i = 0; i < 4; ++i) {
        p.O[i] = E[i];
        p.tl[i] = ceil_div(E[i], p.b[i]);
    }
    p.m_tiles = (long long)p.tl[0] * p.tl[1] * p.tl[2] * p.tl[3];
    const int inner = p.b[0] * p.b[1] * p.b[2];
    // W taps as descriptor start offsets (one load per chunk instead of kw): every 8-row group of the MMA's A operand
    // must be one W run of the box, i.e. b[0] == 8; consecutive groups are then (8 + kw - 1) rows apart (descriptor SBO)
    const char* ew = getenv("ZSV_HALO_WSHIFT");
    p.wshift = p.spatial && kw > 1 && p.b[0] == 8 && !(ew && atoi(ew) == 0);
    const int halo_rows = (p.wshift ? (p.b[0] + kw - 1) * p.b[1] * p.b[2] : inner) * (p.b[3] + p.S - 1);
    p.a_stage_bytes = align1k((uint32_t)halo_rows * 128u);
    p.nchunks = ceil_div(kdim, 64);
    const int tail = kdim - 64 * (p.nchunks - 1);
    p.tail_box = tail <= 16 ? 16 : (tail <= 32 ? 32 : 64);
    const int nmain = p.tail_box == 64 ? p.nchunks : p.nchunks - 1;
    const int ntaps = kt * kh * kw;
    const int cols16 = (cols + 15) & ~15;
    // N tiling: origins step by multiples of 64 (output panels), the last tile takes the remainder
    const int max_nt = (getenv("ZSV_DEBUG_HALO_NSPLIT") && !pair) ? 4 : 1;
    if (pair && p.m_tiles < 2) return p;
    for (int nt = 1; nt <= max_nt; ++nt) {
        int n_step, bn;
        if (nt == 1) {
            n_step = bn = cols16;
        } else {
            n_step = ((cols16 / nt) / 64) * 64;
            if (n_step < 64) break;
            bn = cols16 - n_step * (nt - 1);
            if (bn < n_step) bn = n_step;
        }
        if (bn > 256) continue;
        const uint32_t b_rows = pair ? (uint32_t)bn / 2 : (uint32_t)bn;   // bn is a multiple of 16: whole swizzle atoms
        p.b_main_bytes = align1k(b_rows * 128u);
        p.b_tail_bytes = align1k(b_rows * (uint32_t)p.tail_box * 2u);
        p.b_total_bytes = (uint32_t)ntaps * (nmain * p.b_main_bytes + (nmain < p.nchunks ? p.b_tail_bytes : 0u));
        const int staging1 = ((bn + 63) / 64) * (int)kPanelBytes;
        int nstg = 2, stages = 0, fixed = 0;
        for (; nstg >= 1; --nstg) {   // prefer two staging buffers if >= 3 activation stages still fit
            p.scratch = scratch_mode != 0 ? bn_scratch_bytes(cpad(cols), 0) : 0;
            // (fused BatchNorm backward: one more tile-sized buffer, for y)
            fixed = 1024 + (int)p.b_total_bytes + nstg * staging1 + (scratch_mode == 2 ? staging1 : 0) + p.scratch + 256;
            const int avail = 226 * 1024 - fixed;
            stages = avail > 0 ? avail / (int)p.a_stage_bytes : 0;
            if (stages >= (nstg == 2 ? 3 : 2)) break;
        }
        if (nstg < 1) continue;
        // fused BatchNorm backward: a second y buffer (y requested a whole tile ahead) if the ring stays deep enough and
        // loses at most one stage
        p.nybuf = scratch_mode == 2 ? 1 : 0;
        if (scratch_mode == 2 && !(getenv("ZSV_NYBUF") && atoi(getenv("ZSV_NYBUF")) == 1)) {
            const int avail2 = 226 * 1024 - fixed - staging1;
            const int stages2 = avail2 > 0 ? avail2 / (int)p.a_stage_bytes : 0;
            // (the ring depths plan_halo asks of a pair plan must survive: 3 with the W-shift layout, else 4)
            if (stages2 >= (p.wshift ? 3 : 4) && stages2 >= std::min(stages, 8) - 1) p.nybuf = 2, fixed += staging1, stages = stages2;
        }
        stages = std::min(stages, 8);
        if (const char* e = getenv("ZSV_DEBUG_STAGES")) stages = std::max(2, std::min(stages, atoi(e)));
        if (const char* e = getenv("ZSV_DEBUG_NSTG")) {
            if (atoi(e) == 1 && nstg == 2) nstg = 1;
        }
        p.stages = stages;
        p.nstg = nstg;
        p.bn_tile = bn;
        p.n_step = n_step;
        p.n_tiles = nt;
        p.smem = fixed + stages * (int)p.a_stage_bytes + 16 * stages + 96;
        p.ok = true;
        return p;
    }
    return p;
}

// The pair variant is chosen where the resident weight image starves the activation ring (or does not fit at all):
// at most 2 stages alone, and at least 4 with the image split over the pair.  ZSV_HALO_2CTA=0 / 1 forces never / whenever
// it plans.
HaloPlan plan_halo(int W, int H, int T, int N, int kdim, int cols, int kt, int kh, int kw, int scratch_mode = 0) {
    const HaloPlan one = plan_halo_impl(W, H, T, N, kdim, cols, kt, kh, kw, scratch_mode, false);
    const char* e = getenv("ZSV_HALO_2CTA");
    const int mode = e ? atoi(e) : -1;
    if (mode == 0) return one;
    const bool narrow_temporal = one.ok && !one.spatial && cols <= 64 && kdim > 64;
    if (mode < 0 && one.ok && one.stages > (one.wshift ? 1 : 2) && !narrow_temporal) return one;
    const HaloPlan two = plan_halo_impl(W, H, T, N, kdim, cols, kt, kh, kw, scratch_mode, true);
    if (getenv("ZSV_DEBUG_PLAN"))
        fprintf(stderr, "[zsv] halo plan k=%d cols=%d taps=%dx%dx%d: single ok=%d stages=%d nstg=%d box=%d,%d,%d,%d | pair ok=%d stages=%d nstg=%d\n",
                kdim, cols, kt, kh, kw, (int)one.ok, one.stages, one.nstg, one.b[0], one.b[1], one.b[2], one.b[3], (int)two.ok,
                two.stages, two.nstg);
    if (!two.ok) return one;
    if (mode == 1) return two;
    // narrow temporal convolutions with more than one channel chunk (144->64 fprop): the N = 64 MMAs saturate the
    // shared-memory port (48 clk each, 43 in a pair: profiles/r02_mma_rate.txt) and, since the split epilogue took the
    // statistics off the chain, the pair is faster (105.4 -> 97.9 us at batch 22; 45->64 with one chunk is not)
    if (!one.spatial && one.ok && cols <= 64 && kdim > 64 && two.stages >= 3) return two;
    // measured (profiles/r01_halo_pair_ab.txt): the pair pays off where it reaches a 4-deep ring and the single-CTA
    // plan has at most two stages or does not exist (the generic kernel would run); epilogue-bound temporal
    // convolutions with a 3-stage single-CTA plan are faster as they are
    // (a stage of the W-shift layout carries all kw W taps: three of them are as much work in flight as nine before)
    return two.stages >= (two.wshift ? 3 : 4) ? two : one;
}

// grid of the halo kernel: a multiple of n_tiles so that every CTA keeps one N tile
int halo_grid(const HaloPlan& p) {
    if (p.pair) return 2 * (int)std::min<long long>((p.m_tiles + 1) / 2, sm_count() / 2);
    const long long want = p.m_tiles * p.n_tiles;
    int grid = (int)std::min<long long>(want, (long long)(sm_count() / p.n_tiles) * p.n_tiles);
    if (grid < p.n_tiles) grid = p.n_tiles;
    return grid;
}

// act: tensor [N][T][H][W][pitch] providing the reduction channels (x for fprop, dy for dgrad); wimg: packed weight
// image [tap][rows = output channels][kpitch]; out: [N][T][H][W][opitch].  org / btap describe the tap geometry.
int launch_halo(const HaloPlan& p, const void* act, int actC, int actPitch, const void* wimg, int wRows, int wKpitch,
                int ntaps, void* out, int outPitch, int W, int H, int T, int N, const int* copy_off, int shift_org,
                int tap0, int tap_dcp, int tap_dsh, const void* addend, float* part_sum, float* part_sq, const float* bias,
                int nbias, int relu, cudaStream_t st, BnFuseLaunch* fuse = nullptr) {
    HaloArgs a;
    memset(&a, 0, sizeof(a));
    for (int i = 0; i < 4; ++i) a.b[i] = p.b[i], a.tl[i] = p.tl[i], a.O[i] = p.O[i];
    // strides (elements of the OUTPUT tensor) in box order
    const long long sW = outPitch, sH = (long long)outPitch * W, sT = sH * H, sN = sT * T;
    if (p.spatial) a.os[0] = sW, a.os[1] = sT, a.os[2] = sN, a.os[3] = sH;
    else a.os[0] = sW, a.os[1] = sH, a.os[2] = sN, a.os[3] = sT;
    a.S = p.S, a.ncopies = p.ncopies, a.shift_org = shift_org;
    for (int c = 0; c < p.ncopies; ++c) a.copy_off[c] = copy_off[c];
    a.wshift = p.wshift ? 1 : 0;
    a.w_ext = p.b[0] + (p.wshift ? p.ncopies - 1 : 0);
    a.min_off = copy_off[0];
    for (int c = 1; c < p.ncopies; ++c) a.min_off = std::min(a.min_off, copy_off[c]);
    a.w_first = copy_off[0] - a.min_off;
    a.w_step = p.ncopies > 1 ? copy_off[1] - copy_off[0] : 0;
    for (int c = 2; c < p.ncopies && p.wshift; ++c)
        if (copy_off[c] - copy_off[c - 1] != a.w_step) return fail(ZSV_ERR_UNSUPPORTED, "halo igemm: W taps must be evenly spaced");
    a.tap0 = tap0, a.tap_dcp = tap_dcp, a.tap_dsh = tap_dsh;
    a.kdim = actC, a.nchunks = p.nchunks, a.tail_box = p.tail_box, a.ntaps = ntaps;
    a.ncols = outPitch, a.nbias = nbias, a.bn_tile = p.bn_tile, a.n_step = p.n_step, a.n_tiles = p.n_tiles;
    a.m_tiles = (int)p.m_tiles, a.stages = p.stages, a.relu = relu, a.tmem_cols = 2 * pow2_cols(p.bn_tile);
    a.part_pitch = outPitch;
    a.nstg = p.nstg;
    a.nybuf = fuse ? std::max(p.nybuf, 1) : 0;
    a.fd_tl0 = make_fastdiv(p.tl[0]), a.fd_tl1 = make_fastdiv(p.tl[1]), a.fd_tl2 = make_fastdiv(p.tl[2]);
    if (const char* e = getenv("ZSV_DEBUG_EPI")) a.debug = atoi(e);
    a.pf_dist = 3;
    if (const char* e = getenv("ZSV_DEBUG_PF")) a.pf_dist = atoi(e);
    a.a_stage_bytes = p.a_stage_bytes, a.b_main_bytes = p.b_main_bytes, a.b_tail_bytes = p.b_tail_bytes;
    a.b_total_bytes = p.b_total_bytes;
    a.addend = (const __nv_bfloat16*)addend, a.part_sum = part_sum, a.part_sq = part_sq, a.bias = bias;
    if (a.tmem_cols > 512) return fail(ZSV_ERR_UNSUPPORTED, "halo igemm: N tile too wide");

    // activation maps in box order
    auto act_map = [&](CUtensorMap* m, const void* basep, int C, int pitch, int boxc, int shift_ext,
                       CUtensorMapSwizzle sw, int w_ext) {
        const uint64_t cB = (uint64_t)pitch * 2;
        const uint64_t bW = cB, bH = cB * W, bT = bH * H, bN = bT * T;
        uint64_t dims[5], str[4];
        dims[0] = C;
        dims[1] = W, str[0] = bW;
        if (p.spatial) {
            dims[2] = T, str[1] = bT;
            dims[3] = N, str[2] = bN;
            dims[4] = H, str[3] = bH;
        } else {
            dims[2] = H, str[1] = bH;
            dims[3] = N, str[2] = bN;
            dims[4] = T, str[3] = bT;
        }
        uint32_t box[5] = {(uint32_t)boxc, (uint32_t)w_ext, (uint32_t)p.b[1], (uint32_t)p.b[2], (uint32_t)shift_ext};
        return make_map(m, basep, 5, dims, str, box, sw);
    };
    const CUtensorMapSwizzle tail_sw = p.tail_box == 16 ? CU_TENSOR_MAP_SWIZZLE_32B
                                     : (p.tail_box == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
    CUtensorMap mA, mAt, mB, mBt, mO;
    int rc = act_map(&mA, act, actC, actPitch, 64, p.b[3] + p.S - 1, CU_TENSOR_MAP_SWIZZLE_128B, a.w_ext);
    if (rc) return rc;
    rc = act_map(&mAt, act, actC, actPitch, p.tail_box, p.b[3] + p.S - 1, tail_sw, a.w_ext);
    if (rc) return rc;
    {
        uint64_t dims[3] = {(uint64_t)actC, (uint64_t)wRows, (uint64_t)ntaps};
        uint64_t str[2] = {(uint64_t)wKpitch * 2, (uint64_t)wKpitch * 2 * wRows};
        uint32_t box[3] = {64, (uint32_t)(p.pair ? p.bn_tile / 2 : p.bn_tile), 1};
        rc = make_map(&mB, wimg, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
        box[0] = (uint32_t)p.tail_box;
        rc = make_map(&mBt, wimg, 3, dims, str, box, tail_sw);
        if (rc) return rc;
    }
    rc = act_map(&mO, out, outPitch, outPitch, 64, p.b[3], CU_TENSOR_MAP_SWIZZLE_128B, p.b[0]);
    if (rc) return rc;
    CUtensorMap mY = mO;
    if (fuse) {   // y of the output positions (fused BatchNorm backward): the geometry of the output map
        rc = act_map(&mY, fuse->y, outPitch, outPitch, 64, p.b[3], CU_TENSOR_MAP_SWIZZLE_128B, p.b[0]);
        if (rc) return rc;
    }

    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(igemm_halo_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (attr_err == cudaSuccess)
            attr_err = cudaFuncSetAttribute(igemm_halo_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (attr_err == cudaSuccess)
            attr_err = cudaFuncSetAttribute(igemm_halo_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (attr_err == cudaSuccess)
            attr_err = cudaFuncSetAttribute(igemm_halo_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    if (attr_err != cudaSuccess)
        return fail(ZSV_ERR_CUDA, "cudaFuncSetAttribute(halo igemm) failed: %s", cudaGetErrorString(attr_err));
    const int grid = halo_grid(p);
    a.scratch_bytes = p.scratch;
    if ((fuse || part_sum != nullptr) && p.scratch < bn_scratch_bytes(a.ncols, 0)) {
        return fail(ZSV_ERR_UNSUPPORTED, "halo plan without statistics scratch");
    }
    if (fuse) {
        if (fuse->rows_used + grid > fuse->capacity)
            return fail(ZSV_ERR_WORKSPACE, "dgrad: BN-fusion partial buffer holds %d rows, need %d", fuse->capacity,
                        fuse->rows_used + grid);
        a.bn_y = fuse->y, a.bn_tab = fuse->tab, a.bn_relu = fuse->relu;
        a.bn_partial = fuse->partial + (size_t)fuse->rows_used * 4 * a.ncols;
        fuse->rows_used += grid;
    }
    const bool split = epilogue_split(p.bn_tile, p.nstg);
    if (p.pair) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(grid), cfg.blockDim = dim3(kIgemmThreads), cfg.dynamicSmemBytes = p.smem, cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
        pdl_attribute(&attr[1]);
        cfg.attrs = attr, cfg.numAttrs = 2;
        cudaError_t e = split ? cudaLaunchKernelEx(&cfg, igemm_halo_kernel<true, true>, mA, mAt, mB, mBt, mO, mY, a)
                              : cudaLaunchKernelEx(&cfg, igemm_halo_kernel<true, false>, mA, mAt, mB, mBt, mO, mY, a);
        if (e != cudaSuccess) return fail(ZSV_ERR_CUDA, "launch of igemm_halo_kernel<pair> failed: %s", cudaGetErrorString(e));
    } else if (split) {
        zsv::launch(igemm_halo_kernel<false, true>, grid, kIgemmThreads, p.smem, st, mA, mAt, mB, mBt, mO, mY, a);
    } else {
        zsv::launch(igemm_halo_kernel<false, false>, grid, kIgemmThreads, p.smem, st, mA, mAt, mB, mBt, mO, mY, a);
    }
    ZSV_LAUNCH_CHECK("igemm_halo_kernel");
    return ZSV_OK;
}

}  // namespace
}  // namespace zsv

using namespace zsv;

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" int zsv_conv3d_out_shape(const zsv_conv_desc* d, int32_t out[3]) {
    Shape s;
    int rc = check_desc(d, &s);
    if (rc) return rc;
    out[0] = s.To;
    out[1] = s.Ho;
    out[2] = s.Wo;
    return ZSV_OK;
}

extern "C" size_t zsv_conv3d_packed_weight_bytes(const zsv_conv_desc* d, int which) {
    Shape s;
    if (check_desc(d, &s)) return 0;
    if (which == 0) return (size_t)s.ftaps * d->Cout * s.kpitch * 2;
    if (s.wfold) return 0;
    return (size_t)s.ntaps * d->Cin * s.coutp * 2;
}

extern "C" int zsv_conv3d_pack_weight(const zsv_conv_desc* d, const float* w, void* w_fprop, void* w_dgrad,
                                      void* stream) {
    Shape s;
    int rc = check_desc(d, &s);
    if (rc) return rc;
    if (!w) return fail(ZSV_ERR_BAD_ARG, "pack_weight: null weight");
    if (s.wfold && w_dgrad) return fail(ZSV_ERR_UNSUPPORTED, "pack_weight: no dgrad image for the wfold layout");
    if (!w_fprop && !w_dgrad) return ZSV_OK;
    const long long n = (w_fprop ? (long long)s.ftaps * d->Cout * s.kpitch : 0) +
                        (w_dgrad ? (long long)s.ntaps * d->Cin * s.coutp : 0);
    const int blocks = (int)std::min<long long>(ceil_div_ll(n, 256), 148 * 16);
    zsv::launch(pack_weight_kernel, blocks, 256, 0, (cudaStream_t)stream, w, (__nv_bfloat16*)w_fprop, (__nv_bfloat16*)w_dgrad, d->Cout, d->Cin, s.ntaps, s.kpitch, s.coutp,
        s.wfold ? d->kw : 0);
    ZSV_LAUNCH_CHECK("pack_weight_kernel");
    return ZSV_OK;
}

static int pack_weights_impl(int n, const zsv_conv_desc* descs, const float* const* w, void* const* w_fprop,
                             void* const* w_dgrad, const zsv_bn_fold* fold, void* stream) {
    if (n < 0 || (n > 0 && (!descs || !w || !w_fprop || !w_dgrad)))
        return fail(ZSV_ERR_BAD_ARG, "pack_weights: null array");
    cudaStream_t st = (cudaStream_t)stream;
    const bool tiled_ok = !getenv("ZSV_DEBUG_PACK_ELEMENTWISE");
    for (int base = 0; base < n; base += kMaxPackItems) {
        PackBatch B;
        memset(&B, 0, sizeof(B));
        PackTileBatch TB;
        memset(&TB, 0, sizeof(TB));
        long long total = 0;
        const int cnt = std::min(kMaxPackItems, n - base);
        int m = 0;
        for (int i = 0; i < cnt; ++i) {
            const zsv_conv_desc* d = &descs[base + i];
            Shape s;
            int rc = check_desc(d, &s);
            if (rc) return rc;
            if (!w[base + i]) return fail(ZSV_ERR_BAD_ARG, "pack_weights: null weight %d", base + i);
            if (s.wfold && w_dgrad[base + i])
                return fail(ZSV_ERR_UNSUPPORTED, "pack_weights: no dgrad image for the wfold layout");
            PackItem& I = B.it[m];
            I.w = w[base + i];
            I.wf = (__nv_bfloat16*)w_fprop[base + i];
            I.wd = (__nv_bfloat16*)w_dgrad[base + i];
            I.Cout = d->Cout, I.Cin = d->Cin, I.ntaps = s.ntaps, I.kpitch = s.kpitch, I.copitch = s.coutp;
            I.wfold_kw = s.wfold ? d->kw : 0;
            I.nf = I.wf ? (long long)s.ftaps * d->Cout * s.kpitch : 0;
            I.nd = I.wd ? (long long)s.ntaps * d->Cin * s.coutp : 0;
            long long nb = 0;
            if (fold && fold[base + i].running_var) {
                const zsv_bn_fold& f = fold[base + i];
                if (!f.running_mean || !f.bias_out) return fail(ZSV_ERR_BAD_ARG, "pack_weights: incomplete zsv_bn_fold %d", base + i);
                I.gamma = f.gamma, I.beta = f.beta, I.rmean = f.running_mean, I.rvar = f.running_var;
                I.bias_out = f.bias_out, I.eps = f.eps;
                nb = d->Cout;
            }
            if (I.nf + I.nd + nb == 0) continue;
            if (tiled_ok && !s.wfold && 8 * s.ntaps <= kPackElems) {
                // plain layout: the tiled kernel (this slot of the element-wise batch is reused by the next item)
                PackTileItem& T = TB.it[TB.n];
                T.w = I.w, T.gamma = I.gamma, T.beta = I.beta, T.rmean = I.rmean, T.rvar = I.rvar, T.bias_out = I.bias_out;
                T.wf = I.wf, T.wd = I.wd, T.eps = I.eps;
                T.Cout = I.Cout, T.Cin = I.Cin, T.ntaps = I.ntaps, T.kpitch = I.kpitch, T.copitch = I.copitch;
                T.ci_chunk = std::max(8, std::min(64, (kPackElems / s.ntaps) & ~7));
                T.n_ci = ceil_div(d->Cin, T.ci_chunk);
                T.unit_start = TB.units;
                TB.units += ceil_div(T.wd ? s.coutp : d->Cout, kPackCo) * T.n_ci;
                ++TB.n;
                memset(&I, 0, sizeof(I));
                continue;
            }
            I.start = total;
            total += I.nf + I.nd + nb;
            ++m;
        }
        if (TB.n > 0) {
            zsv::launch(pack_weights_tiled_kernel, TB.units, 256, 0, st, TB);
            ZSV_LAUNCH_CHECK("pack_weights_tiled_kernel");
        }
        if (m == 0) continue;
        B.n = m;
        B.total = total;
        const int blocks = (int)std::min<long long>(ceil_div_ll(total, 256), (long long)sm_count() * 16);
        zsv::launch(pack_weights_batched_kernel, blocks, 256, 0, st, B);
        ZSV_LAUNCH_CHECK("pack_weights_batched_kernel");
    }
    return ZSV_OK;
}

extern "C" int zsv_conv3d_pack_weights(int n, const zsv_conv_desc* descs, const float* const* w, void* const* w_fprop,
                                       void* const* w_dgrad, void* stream) {
    return pack_weights_impl(n, descs, w, w_fprop, w_dgrad, nullptr, stream);
}

extern "C" int zsv_conv3d_pack_weights_folded(int n, const zsv_conv_desc* descs, const float* const* w,
                                              void* const* w_fprop, const zsv_bn_fold* fold, void* stream) {
    if (n > 0 && !fold) return fail(ZSV_ERR_BAD_ARG, "pack_weights_folded: null fold array");
    std::vector<void*> none((size_t)std::max(n, 1), nullptr);
    return pack_weights_impl(n, descs, w, w_fprop, none.data(), fold, stream);
}

extern "C" int zsv_conv3d_stat_rows(const zsv_conv_desc* d) {
    Shape s;
    if (check_desc(d, &s)) return -1;
    // one partial row per CTA of the (persistent) fprop kernel; which kernel runs depends on the geometry
    if (!s.wfold && d->st == 1 && d->sh == 1 && d->sw == 1 && s.To == d->T && s.Ho == d->H && s.Wo == d->W) {
        const HaloPlan hp = plan_halo(d->W, d->H, d->T, d->N, d->Cin, d->Cout, d->kt, d->kh, d->kw, 1);
        if (hp.ok) return halo_grid(hp);
    }
    const Box b = choose_box(s.Wo, s.Ho, s.To, d->N, false);
    const long long m_tiles = (long long)ceil_div(s.Wo, b.bw) * ceil_div(s.Ho, b.bh) * ceil_div(s.To, b.bt) * ceil_div(d->N, b.bn);
    int bn_tile, n_tiles;
    choose_ntile(d->Cout, m_tiles, &bn_tile, &n_tiles);
    return igemm_grid(bn_tile, m_tiles, n_tiles);
}

extern "C" int zsv_conv3d_fprop(const zsv_conv_desc* d, const void* x, const void* w_fprop, void* y, float* part_sum,
                                float* part_sq, const float* bias, const void* addend, int relu, void* stream) {
    Shape s;
    int rc = check_desc(d, &s);
    if (rc) return rc;
    if (!x || !w_fprop || !y) return fail(ZSV_ERR_BAD_ARG, "fprop: null pointer");
    if ((part_sum == nullptr) != (part_sq == nullptr)) return fail(ZSV_ERR_BAD_ARG, "fprop: need both stat buffers");

    if (!s.wfold && d->st == 1 && d->sh == 1 && d->sw == 1 && s.To == d->T && s.Ho == d->H && s.Wo == d->W) {
        const HaloPlan hp = plan_halo(d->W, d->H, d->T, d->N, d->Cin, d->Cout, d->kt, d->kh, d->kw, part_sum ? 1 : 0);
        if (hp.ok) {
            // tap(cp, sh) = sh*kw + cp (spatial) or sh (temporal)
            int copy_off[kMaxCopies];
            for (int c = 0; c < hp.ncopies; ++c) copy_off[c] = hp.spatial ? c - d->pw : 0;
            const int shift_org = hp.spatial ? -d->ph : -d->pt;
            return launch_halo(hp, x, d->Cin, s.cinp, w_fprop, d->Cout, s.kpitch, s.ntaps, y, s.coutp, d->W, d->H, d->T,
                               d->N, copy_off, shift_org, 0, 1, hp.spatial ? d->kw : 1, addend, part_sum, part_sq, bias,
                               bias ? d->Cout : 0, relu, (cudaStream_t)stream);
        }
    }

    IgemmArgs a;
    memset(&a, 0, sizeof(a));
    const Box b = choose_box(s.Wo, s.Ho, s.To, d->N, false);
    a.bw = b.bw, a.bh = b.bh, a.bt = b.bt, a.bn = b.bn;
    a.tw = ceil_div(s.Wo, b.bw), a.th = ceil_div(s.Ho, b.bh), a.tt = ceil_div(s.To, b.bt), a.tn = ceil_div(d->N, b.bn);
    a.OW = s.Wo, a.OH = s.Ho, a.OT = s.To, a.ON = d->N;
    a.kdim = s.keff;
    a.ntaps = s.ftaps;
    a.ncols = s.coutp;
    a.nbias = bias ? d->Cout : 0;
    a.relu = relu;
    a.part_pitch = s.coutp;
    a.o_sW = s.coutp;
    a.o_sH = (long long)s.coutp * s.Wo;
    a.o_sT = a.o_sH * s.Ho;
    a.o_sN = a.o_sT * s.To;
    a.out = (__nv_bfloat16*)y;
    a.addend = (const __nv_bfloat16*)addend;
    a.part_sum = part_sum;
    a.part_sq = part_sq;
    a.bias = bias;
    const long long m_tiles = (long long)a.tw * a.th * a.tt * a.tn;
    int n_tiles;
    choose_ntile(d->Cout, m_tiles, &a.bn_tile, &n_tiles);

    CUtensorMap maps[kMaxMaps];
    int nmaps;
    rc = build_fwd_taps(d, s, x, b, a.taps, maps, &nmaps);
    if (rc) return rc;
    CUtensorMap mapB;
    {
        uint64_t dims[3] = {(uint64_t)s.keff, (uint64_t)d->Cout, (uint64_t)s.ftaps};
        uint64_t strides[2] = {(uint64_t)s.kpitch * 2, (uint64_t)s.kpitch * 2 * d->Cout};
        // a CTA pair loads half of the B rows per CTA
        uint32_t box[3] = {64, (uint32_t)(igemm_use_pair(a.bn_tile, m_tiles) ? a.bn_tile / 2 : a.bn_tile), 1};
        rc = make_map(&mapB, w_fprop, 3, dims, strides, box);
        if (rc) return rc;
    }
    CUtensorMap mapOut;
    rc = make_plain_map(&mapOut, y, d->N, s.To, s.Ho, s.Wo, s.coutp, s.coutp, b);
    if (rc) return rc;
    return launch_igemm(maps, mapB, mapOut, a, m_tiles, n_tiles, (cudaStream_t)stream);
}

extern "C" int zsv_conv3d_dgrad(const zsv_conv_desc* d, const void* dy, const void* w_dgrad, void* dx,
                                const void* addend, zsv_bn_bwd_fuse* bnf, void* stream) {
    Shape s;
    int rc = check_desc(d, &s);
    if (rc) return rc;
    if (s.wfold) return fail(ZSV_ERR_UNSUPPORTED, "dgrad: not available for the wfold (first-layer) layout");
    if (!dy || !w_dgrad || !dx) return fail(ZSV_ERR_BAD_ARG, "dgrad: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    BnFuseLaunch fuse_state;
    BnFuseLaunch* fuse = nullptr;
    if (bnf) {
        if (!bnf->y || !bnf->table || !bnf->partial || bnf->partial_rows < 1)
            return fail(ZSV_ERR_BAD_ARG, "dgrad: incomplete zsv_bn_bwd_fuse");
        fuse_state.y = (const __nv_bfloat16*)bnf->y;
        fuse_state.tab = (const float4*)bnf->table;
        fuse_state.relu = bnf->relu;
        fuse_state.partial = bnf->partial;
        fuse_state.capacity = bnf->partial_rows;
        fuse_state.rows_used = 0;
        fuse = &fuse_state;
        bnf->rows_written = 0;
    }

    if (d->st == 1 && d->sh == 1 && d->sw == 1 && s.To == d->T && s.Ho == d->H && s.Wo == d->W) {
        const HaloPlan hp = plan_halo(d->W, d->H, d->T, d->N, d->Cout, d->Cin, d->kt, d->kh, d->kw, fuse != nullptr ? 2 : 0);
        if (hp.ok) {
            // dx[i] = sum_j dy[i + p - j] w[j]: copy j reads dy at W offset p - j; along the shift dim the halo starts
            // at i0 + p - (S-1) and filter index j sits at halo slice S-1-j
            // (tap = (S-1-sh)*kw + cp for spatial, S-1-sh for temporal)
            int copy_off[kMaxCopies];
            for (int c = 0; c < hp.ncopies; ++c) copy_off[c] = hp.spatial ? d->pw - c : 0;
            const int shift_org = (hp.spatial ? d->ph : d->pt) - (hp.S - 1);
            const int kwm = hp.spatial ? d->kw : 1;
            rc = launch_halo(hp, dy, d->Cout, s.coutp, w_dgrad, d->Cin, s.coutp, s.ntaps, dx, s.cinp, d->W, d->H, d->T,
                             d->N, copy_off, shift_org, (hp.S - 1) * kwm, 1, -kwm, addend, nullptr, nullptr, nullptr, 0, 0,
                             st, fuse);
            if (fuse) bnf->rows_written = fuse->rows_used;
            return rc;
        }
    }

    // weight image [tap][Cin][coutp]: K = Cout
    const int classes_t = d->st, classes_h = d->sh, classes_w = d->sw;
    bool any_empty = false;
    for (int ct = 0; ct < classes_t; ++ct)
        for (int ch = 0; ch < classes_h; ++ch)
            for (int cw = 0; cw < classes_w; ++cw)
                if (bwd_dim_taps(d->kt, d->st, d->pt, ct).empty() || bwd_dim_taps(d->kh, d->sh, d->ph, ch).empty() ||
                    bwd_dim_taps(d->kw, d->sw, d->pw, cw).empty())
                    any_empty = true;
    const size_t dx_bytes = (size_t)d->N * d->T * d->H * d->W * s.cinp * 2;
    if (any_empty && fuse)
        return fail(ZSV_ERR_UNSUPPORTED, "dgrad: BN-backward fusion needs every output parity class to have a filter tap");
    if (any_empty) {
        if (addend) {
            if (addend != dx) ZSV_CUDA_CHECK(cudaMemcpyAsync(dx, addend, dx_bytes, cudaMemcpyDeviceToDevice, st));
        } else {
            ZSV_CUDA_CHECK(cudaMemsetAsync(dx, 0, dx_bytes, st));
        }
    }

    for (int ct = 0; ct < classes_t; ++ct)
        for (int ch = 0; ch < classes_h; ++ch)
            for (int cw = 0; cw < classes_w; ++cw) {
                std::vector<DimTap> tt = bwd_dim_taps(d->kt, d->st, d->pt, ct);
                std::vector<DimTap> th = bwd_dim_taps(d->kh, d->sh, d->ph, ch);
                std::vector<DimTap> tw = bwd_dim_taps(d->kw, d->sw, d->pw, cw);
                if (tt.empty() || th.empty() || tw.empty()) continue;
                const int QT = (d->T - ct + d->st - 1) / d->st;
                const int QH = (d->H - ch + d->sh - 1) / d->sh;
                const int QW = (d->W - cw + d->sw - 1) / d->sw;
                if (QT < 1 || QH < 1 || QW < 1) continue;
                IgemmArgs a;
                memset(&a, 0, sizeof(a));
                const Box b = choose_box(QW, QH, QT, d->N, false);
                a.bw = b.bw, a.bh = b.bh, a.bt = b.bt, a.bn = b.bn;
                a.tw = ceil_div(QW, b.bw), a.th = ceil_div(QH, b.bh), a.tt = ceil_div(QT, b.bt),
                a.tn = ceil_div(d->N, b.bn);
                a.OW = QW, a.OH = QH, a.OT = QT, a.ON = d->N;
                a.kdim = d->Cout;
                a.ntaps = (int)(tt.size() * th.size() * tw.size());
                a.ncols = s.cinp;
                a.o_sW = (long long)s.cinp * d->sw;
                a.o_sH = (long long)s.cinp * d->W * d->sh;
                a.o_sT = (long long)s.cinp * d->W * d->H * d->st;
                a.o_sN = (long long)s.cinp * d->W * d->H * d->T;
                const long long class_off = (((long long)ct * d->H + ch) * d->W + cw) * s.cinp;
                a.out = (__nv_bfloat16*)dx + class_off;
                // when the tensor was pre-filled with the addend (empty classes), it is still added here for
                // the non-empty classes because their positions are overwritten
                a.addend = addend ? (const __nv_bfloat16*)addend + class_off : nullptr;
                if (fuse) fuse->y = (const __nv_bfloat16*)bnf->y + class_off;
                int ti = 0;
                for (const DimTap& x1 : tt)
                    for (const DimTap& x2 : th)
                        for (const DimTap& x3 : tw) {
                            a.taps[ti].map = 0;
                            a.taps[ti].dt = (int16_t)x1.off;
                            a.taps[ti].dh = (int16_t)x2.off;
                            a.taps[ti].dw = (int16_t)x3.off;
                            a.taps[ti].btap = (int16_t)((x1.j * d->kh + x2.j) * d->kw + x3.j);
                            ++ti;
                        }
                const long long m_tiles = (long long)a.tw * a.th * a.tt * a.tn;
                int n_tiles;
                // fused BatchNorm backward: a second tile-sized buffer (y) lives beside the output staging, so N tiles stay
                // at 192 columns (3 panels) and the ring keeps 3-4 stages
                choose_ntile(d->Cin, m_tiles, &a.bn_tile, &n_tiles, fuse ? 192 : 256);
                CUtensorMap maps[kMaxMaps];
                rc = make_plain_map(&maps[0], dy, d->N, s.To, s.Ho, s.Wo, d->Cout, s.coutp, b);
                if (rc) return rc;
                for (int i = 1; i < kMaxMaps; ++i) maps[i] = maps[0];
                CUtensorMap mapB;
                uint64_t dims[3] = {(uint64_t)d->Cout, (uint64_t)d->Cin, (uint64_t)s.ntaps};
                uint64_t strides[2] = {(uint64_t)s.coutp * 2, (uint64_t)s.coutp * 2 * d->Cin};
                uint32_t box[3] = {64, (uint32_t)(igemm_use_pair(a.bn_tile, m_tiles) ? a.bn_tile / 2 : a.bn_tile), 1};
                rc = make_map(&mapB, w_dgrad, 3, dims, strides, box);
                if (rc) return rc;
                CUtensorMap mapOut;
                {
                    const uint64_t cB = (uint64_t)s.cinp * 2;
                    uint64_t odims[5] = {(uint64_t)s.cinp, (uint64_t)QW, (uint64_t)QH, (uint64_t)QT, (uint64_t)d->N};
                    uint64_t ostr[4] = {cB * d->sw, cB * d->W * d->sh, cB * d->W * d->H * d->st,
                                        cB * d->W * d->H * d->T};
                    uint32_t obox[5] = {64, (uint32_t)b.bw, (uint32_t)b.bh, (uint32_t)b.bt, (uint32_t)b.bn};
                    rc = make_map(&mapOut, (const char*)dx + class_off * 2, 5, odims, ostr, obox);
                    if (rc) return rc;
                }
                CUtensorMap mapY;
                if (fuse) {   // y of the positions of this parity class: the geometry of the output map
                    const uint64_t cB = (uint64_t)s.cinp * 2;
                    uint64_t odims[5] = {(uint64_t)s.cinp, (uint64_t)QW, (uint64_t)QH, (uint64_t)QT, (uint64_t)d->N};
                    uint64_t ostr[4] = {cB * d->sw, cB * d->W * d->sh, cB * d->W * d->H * d->st,
                                        cB * d->W * d->H * d->T};
                    uint32_t obox[5] = {64, (uint32_t)b.bw, (uint32_t)b.bh, (uint32_t)b.bt, (uint32_t)b.bn};
                    rc = make_map(&mapY, fuse->y, 5, odims, ostr, obox);
                    if (rc) return rc;
                }
                rc = launch_igemm(maps, mapB, mapOut, a, m_tiles, n_tiles, st, fuse, fuse ? &mapY : nullptr);
                if (rc) return rc;
            }
    if (fuse) bnf->rows_written = fuse->rows_used;
    return ZSV_OK;
}

namespace {
struct WgradPlan {
    Box box;
    int num_kb, kchunks, npanels, m_tiles, bn_tile, n_tiles, nbp, splits, kb_per_split, stages, ci_pitch, co_pitch, mt;
    size_t ws_bytes;
};
int plan_wgrad(const zsv_conv_desc* d, const Shape& s, WgradPlan* p) {
    p->box = choose_box(s.Wo, s.Ho, s.To, d->N, true);
    const Box& b = p->box;
    p->num_kb = ceil_div(s.Wo, b.bw) * ceil_div(s.Ho, b.bh) * ceil_div(s.To, b.bt) * ceil_div(d->N, b.bn);
    p->kchunks = ceil_div(s.keff, 64);
    p->npanels = s.ftaps * p->kchunks;
    p->m_tiles = ceil_div(p->npanels, 2);
    // N tiles are whole 64-wide dy panels except the last
    const int cols16 = (d->Cout + 15) & ~15;
    p->n_tiles = ceil_div(cols16, 256);
    p->bn_tile = (ceil_div(cols16, p->n_tiles) + 15) & ~15;
    if (p->n_tiles > 1) p->bn_tile = (p->bn_tile + 63) & ~63;  // keep panel-aligned tile origins
    p->n_tiles = ceil_div(cols16, p->bn_tile);
    p->nbp = ceil_div(p->bn_tile, 64);
    // (two M tiles per CTA sharing the dy panels of a stage were measured in round 1: 10 % faster alone on the layer-1 / 2
    // spatial weight gradients, 0.8 % slower inside the training step; the option is gone, the kernel keeps one)
    const int sms = std::max(1, sm_count());
    p->mt = 1;
    p->m_tiles = ceil_div(p->m_tiles, p->mt);
    const int tiles = p->m_tiles * p->n_tiles;
    // split the position axis so that the CTAs fill the GPU ONCE: measured against two waves (the previous choice) the
    // step is 2.8 % faster (13.84 vs 14.23 ms) -- one pipeline fill/drain and one TMEM allocation per SM instead of two,
    // and half as many split partials to write and to reduce in the finalize kernel
    int waves = 1;
    if (const char* e = getenv("ZSV_DEBUG_WGRAD_WAVES")) waves = std::max(1, atoi(e));
    int splits = std::max(1, (waves * sms) / tiles);
    splits = std::min(splits, p->num_kb);
    p->kb_per_split = ceil_div(p->num_kb, splits);
    p->splits = ceil_div(p->num_kb, p->kb_per_split);
    const int stage = (2 * p->mt + p->nbp) * kPanelBytes;
    int smem_cap = 227 * 1024;
    if (const char* e = getenv("ZSV_DEBUG_WGRAD_SMEM_KB")) smem_cap = std::min(smem_cap, std::max(64, atoi(e)) * 1024);
    p->stages = std::max(1, std::min(4, (smem_cap - 2048) / stage));
    p->ci_pitch = s.wfold ? 64 : s.cinp;
    p->co_pitch = p->n_tiles * p->bn_tile;
    p->ws_bytes = (size_t)p->splits * s.ftaps * p->ci_pitch * p->co_pitch * 4;
    return ZSV_OK;
}
}  // namespace

namespace {
// ---- CTA-pair weight gradient (wgrad_pair_kernel) ---------------------------------------------------
struct WgradPairPlan {
    bool ok;
    Box box;
    int rows, slot_rows, num_kb, kchunks, npanels, m_tiles, bn_tile, half_n, nbh, n_tiles, splits, kb_per_split, stages;
    int ci_pitch, co_pitch, tmem_cols, smem;
    size_t ws_bytes;
};

// Plain-layout convolutions with at least one full 256-row M tile (taps x 64-channel chunks >= 4) and >= 128 output
// channels: layers 2-4.  Picks the position box (least padding of the K = 16 granule), the N tiling and the split-K
// factor from a small time model: waves x (k-blocks per item x L2-bound block time + fixed cost per item) + the traffic
// of the split partials.
WgradPairPlan plan_wgrad_pair(const zsv_conv_desc* d, const Shape& s) {
    WgradPairPlan p;
    memset(&p, 0, sizeof(p));
    const char* e = getenv("ZSV_WGRAD_PAIR");
    if (e && atoi(e) == 0) return p;
    if (s.wfold) return p;
    p.kchunks = ceil_div(s.keff, 64);
    p.npanels = s.ftaps * p.kchunks;
    const bool force = e && atoi(e) == 2;      // tests: every layout the kernel can run
    if (!force && (p.npanels < 4 || d->Cout < 128)) return p;
    const int OW = s.Wo, OH = s.Ho, OT = s.To, ON = d->N;
    const double total = (double)OW * OH * OT * ON;
    const int pairs = std::max(1, sm_count() / 2);
    const int cols16 = (d->Cout + 15) & ~15;
    p.m_tiles = ceil_div(p.npanels, 4);
    double best = 1e300;
    for (int nt = ceil_div(cols16, 256); nt <= ceil_div(cols16, 256) + 3 && nt <= ceil_div(cols16, 16); ++nt) {
        const int bn = (ceil_div(cols16, nt) + 15) & ~15;
        if (bn > 256 || (nt > 1 && bn * (nt - 1) >= cols16)) continue;
        const int nbh = ceil_div(bn / 2, 64);
        // widest slot that still leaves a 3-deep ring
        const int max_slot = std::min(256, ((220 * 1024 / 3) / ((2 + nbh) * 128)) & ~15);
        Box bb{1, 1, 1, 1};
        double bscore = -1;
        for (int bw = 1; bw <= std::min(OW, max_slot); ++bw)
            for (int bh = 1; bh <= std::min(OH, max_slot / bw); ++bh)
                for (int bt = 1; bt <= std::min(OT, max_slot / (bw * bh)); ++bt)
                    for (int bnn = 1; bnn <= std::min(ON, max_slot / (bw * bh * bt)); ++bnn) {
                        const int rows = bw * bh * bt * bnn;
                        const int slot = (rows + 15) & ~15;
                        if (slot > max_slot) continue;
                        const double tiles = (double)ceil_div(OW, bw) * ceil_div(OH, bh) * ceil_div(OT, bt) * ceil_div(ON, bnn);
                        // MMA rows spent per useful position, plus ~1.5 k-steps of barrier / issue latency per block
                        const double score = total / (tiles * (slot + 24.0)) + 1e-7 * bw;
                        if (score > bscore) bscore = score, bb = Box{bw, bh, bt, bnn};
                    }
        if (bscore < 0) continue;
        const int rows = bb.rows(), slot = (rows + 15) & ~15;
        const int num_kb = ceil_div(OW, bb.bw) * ceil_div(OH, bb.bh) * ceil_div(OT, bb.bt) * ceil_div(ON, bb.bn);
        // time of one k-block in clocks: L2 -> SM ingest (~42 B/clk/SM) against the MMA itself (4096 MAC/clk/SM)
        const double t_kb = std::max(slot * (2.0 + nbh) * 128.0 / 42.0, slot * 256.0 * bn / 8192.0) + 150.0;
        const long long tiles = (long long)p.m_tiles * nt;
        const double ws_clk = (double)s.ftaps * (double)(s.cinp) * (double)(nt * bn) * 8.0 / 3400.0;   // write + read per split
        for (int sp = 1; sp <= std::min(num_kb, 64); ++sp) {
            const int kbs = ceil_div(num_kb, sp);
            const int sp_eff = ceil_div(num_kb, kbs);
            const double waves = (double)ceil_div_ll(tiles * sp_eff, pairs);
            const double cost = waves * (kbs * t_kb + 4000.0) + sp_eff * ws_clk;
            if (cost < best) {
                best = cost;
                p.box = bb, p.rows = rows, p.slot_rows = slot, p.num_kb = num_kb;
                p.bn_tile = bn, p.half_n = bn / 2, p.nbh = nbh, p.n_tiles = nt;
                p.kb_per_split = kbs, p.splits = sp_eff;
            }
        }
    }
    if (best >= 1e300) return p;
    if (const char* f = getenv("ZSV_DEBUG_WGRAD_SPLITS")) {
        const int sp = std::max(1, std::min(p.num_kb, atoi(f)));
        p.kb_per_split = ceil_div(p.num_kb, sp);
        p.splits = ceil_div(p.num_kb, p.kb_per_split);
    }
    const int stage = (2 + p.nbh) * p.slot_rows * 128;
    p.stages = std::max(2, std::min(6, (225 * 1024 - 1024) / stage));
    p.smem = 1024 + p.stages * stage + 16 * p.stages + 64;
    if (p.smem > 227 * 1024) return p;
    p.tmem_cols = pow2_cols(p.bn_tile);
    p.ci_pitch = s.cinp;
    p.co_pitch = p.n_tiles * p.bn_tile;
    p.ws_bytes = (size_t)p.splits * s.ftaps * p.ci_pitch * p.co_pitch * 4;
    p.ok = true;
    if (getenv("ZSV_DEBUG_PLAN"))
        fprintf(stderr, "[zsv] wgrad pair plan %d->%d taps %d pos %dx%dx%dx%d: box %d,%d,%d,%d rows %d/%d kb %d  M tiles %d  N %d x %d  "
                        "splits %d x %d kb  stages %d\n", d->Cin, d->Cout, s.ftaps, OW, OH, OT, ON, p.box.bw, p.box.bh, p.box.bt,
                p.box.bn, p.rows, p.slot_rows, p.num_kb, p.m_tiles, p.n_tiles, p.bn_tile, p.splits, p.kb_per_split, p.stages);
    return p;
}

int launch_wgrad_pair(const WgradPairPlan& p, const zsv_conv_desc* d, const Shape& s, const void* x, const void* dy,
                      void* workspace, cudaStream_t st) {
    WgradPairArgs a;
    memset(&a, 0, sizeof(a));
    const Box& b = p.box;
    a.bw = b.bw, a.bh = b.bh, a.bt = b.bt, a.bn = b.bn;
    a.tw = ceil_div(s.Wo, b.bw), a.th = ceil_div(s.Ho, b.bh), a.tt = ceil_div(s.To, b.bt), a.tn = ceil_div(d->N, b.bn);
    a.rows = p.rows, a.slot_rows = p.slot_rows;
    a.kchunks = p.kchunks, a.npanels = p.npanels, a.ntaps = s.ftaps;
    a.ci_store = p.ci_pitch, a.ci_pitch = p.ci_pitch, a.co_pitch = p.co_pitch;
    a.bn_tile = p.bn_tile, a.half_n = p.half_n, a.nbh = p.nbh;
    a.stages = p.stages, a.tmem_cols = p.tmem_cols;
    a.num_kb = p.num_kb, a.kb_per_split = p.kb_per_split;
    a.ws = (float*)workspace;
    CUtensorMap maps[kMaxMaps];
    int nmaps;
    int rc = build_fwd_taps(d, s, x, b, a.taps, maps, &nmaps);
    if (rc) return rc;
    CUtensorMap mapB;
    rc = make_plain_map(&mapB, dy, d->N, s.To, s.Ho, s.Wo, d->Cout, s.coutp, b);
    if (rc) return rc;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(wgrad_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    if (attr_err != cudaSuccess)
        return fail(ZSV_ERR_CUDA, "cudaFuncSetAttribute(wgrad pair) failed: %s", cudaGetErrorString(attr_err));
    MapPack pack;
    for (int i = 0; i < kMaxMaps; ++i) pack.m[i] = maps[i];
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * p.m_tiles, p.n_tiles, p.splits), cfg.blockDim = dim3(192), cfg.dynamicSmemBytes = p.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    pdl_attribute(&attr[1]);
    cfg.attrs = attr, cfg.numAttrs = 2;
    cudaError_t e = cudaLaunchKernelEx(&cfg, wgrad_pair_kernel, pack, mapB, a);
    if (e != cudaSuccess) return fail(ZSV_ERR_CUDA, "launch of wgrad_pair_kernel failed: %s", cudaGetErrorString(e));
    ZSV_LAUNCH_CHECK("wgrad_pair_kernel");
    return ZSV_OK;
}

struct WgradHaloPlan {
    bool ok, spatial;
    int pair_taps, nacc, ncopies;
    int b[4], tl[4];
    int nchunks, npairs, bn_tile, nbp, stages, acc_stride, tmem_cols, num_kb, kb_per_split, splits, ci_pitch, co_pitch;
    uint32_t a_chunk_bytes, stage_bytes;
    int smem;
    size_t ws_bytes;
};

// Stride-1 "same" convolutions on plain NDHWC input whose accumulators fit 512 TMEM columns:
//   temporal 3x1x1: halo along T, one CTA column;  spatial 1x3xkw: halo along H, one CTA column per filter column.
WgradHaloPlan plan_wgrad_halo(const zsv_conv_desc* d, const Shape& s) {
    WgradHaloPlan p;
    memset(&p, 0, sizeof(p));
    if (getenv("ZSV_DEBUG_NO_WGRAD_HALO")) return p;
    if (s.wfold || d->st != 1 || d->sh != 1 || d->sw != 1) return p;
    const bool temporal = d->kt == 3 && d->kh == 1 && d->kw == 1 && d->pt == 1 && d->ph == 0 && d->pw == 0;
    const bool spatial = d->kt == 1 && d->kh == 3 && d->kw >= 1 && d->kw <= 3 && d->pt == 0 && d->ph == 1 &&
                         d->pw == d->kw / 2 && !getenv("ZSV_DEBUG_NO_WGRAD_HALO_SPATIAL");
    if (!temporal && !spatial) return p;
    p.spatial = spatial;
    const int cols16 = (d->Cout + 15) & ~15;
    if (cols16 > 256) return p;
    p.bn_tile = cols16;
    p.nbp = ceil_div(p.bn_tile, 64);
    p.nchunks = ceil_div(d->Cin, 64);
    p.npairs = ceil_div(p.nchunks, 2);
    p.acc_stride = (p.bn_tile + 31) & ~31;
    p.pair_taps = p.nchunks == 1;
    p.nacc = p.pair_taps ? 2 : 3 * p.npairs;          // 3 halo taps
    if (p.nacc * p.acc_stride > 512) return p;
    p.tmem_cols = pow2_cols(p.nacc * p.acc_stride);
    p.ncopies = spatial ? d->kw : 1;
    // box: inner dims then the halo dim -- temporal (W, H, N | T), spatial (W, T, N | H)
    const int E[4] = {d->W, spatial ? d->T : d->H, d->N, spatial ? d->H : d->T};
    double best = -1;
    for (int b0 = 1; b0 <= std::min(E[0], 128); ++b0)
        for (int b1 = 1; b1 <= std::min(E[1], 128 / b0); ++b1)
            for (int b2 = 1; b2 <= std::min(E[2], 128 / (b0 * b1)); ++b2) {
                const int inner = b0 * b1 * b2;
                if (inner % 8) continue;
                for (int b3 = 1; b3 <= std::min(E[3], 128 / inner); ++b3) {
                    if ((inner * b3) % 16) continue;
                    const double tiles = (double)ceil_div(E[0], b0) * ceil_div(E[1], b1) * ceil_div(E[2], b2) * ceil_div(E[3], b3);
                    const double eff = ((double)E[0] * E[1] * E[2] * E[3]) / (tiles * inner * b3);
                    const double halo = (double)(b3 + 2) / b3;
                    const double score = eff * (inner * b3 / 128.0 + 1.0) / (0.5 + 0.5 * halo) + 1e-6 * b0;
                    if (score > best) {
                        best = score;
                        p.b[0] = b0, p.b[1] = b1, p.b[2] = b2, p.b[3] = b3;
                    }
                }
            }
    if (best < 0) return p;
    for (int i = 0; i < 4; ++i) p.tl[i] = ceil_div(E[i], p.b[i]);
    p.num_kb = p.tl[0] * p.tl[1] * p.tl[2] * p.tl[3];
    const int inner = p.b[0] * p.b[1] * p.b[2];
    // pair_taps reads up to one slice past the halo (the dummy second half of the last accumulator)
    p.a_chunk_bytes = align1k((uint32_t)(inner * (p.b[3] + 2 + (p.pair_taps ? 1 : 0))) * 128u);
    p.stage_bytes = (p.pair_taps ? 1u : 2u * p.npairs) * p.a_chunk_bytes + (uint32_t)p.nbp * kPanelBytes;
    p.stages = std::min(4, (int)((227 * 1024 - 2048) / p.stage_bytes));
    if (p.stages < 2) return p;
    const int splits = std::min(p.num_kb, std::max(1, sm_count() / p.ncopies));
    p.kb_per_split = ceil_div(p.num_kb, splits);
    p.splits = ceil_div(p.num_kb, p.kb_per_split);
    p.ci_pitch = s.cinp;
    p.co_pitch = p.bn_tile;
    p.ws_bytes = (size_t)p.splits * s.ntaps * p.ci_pitch * p.co_pitch * 4;
    p.smem = 1024 + p.stages * (int)p.stage_bytes + 16 * p.stages + 64;
    p.ok = true;
    return p;
}

int launch_wgrad_halo(const WgradHaloPlan& p, const zsv_conv_desc* d, const Shape& s, const void* x, const void* dy,
                      void* workspace, cudaStream_t st) {
    WgradHaloArgs a;
    memset(&a, 0, sizeof(a));
    for (int i = 0; i < 4; ++i) a.b[i] = p.b[i], a.tl[i] = p.tl[i];
    a.fd_tl0 = make_fastdiv(p.tl[0]), a.fd_tl1 = make_fastdiv(p.tl[1]), a.fd_tl2 = make_fastdiv(p.tl[2]);
    a.nchunks = p.nchunks, a.npairs = p.npairs, a.ntaps = 3;
    a.ci_pitch = p.ci_pitch, a.co_pitch = p.co_pitch, a.bn_tile = p.bn_tile, a.nbp = p.nbp;
    a.stages = p.stages, a.tmem_cols = p.tmem_cols, a.acc_stride = p.acc_stride;
    a.num_kb = p.num_kb, a.kb_per_split = p.kb_per_split;
    a.a_chunk_bytes = p.a_chunk_bytes, a.stage_bytes = p.stage_bytes;
    a.ncopies = p.ncopies, a.copy_org = p.spatial ? -d->pw : 0, a.tap_stride = p.spatial ? d->kw : 1;
    a.pair_taps = p.pair_taps, a.nacc = p.nacc, a.total_taps = s.ntaps;
    a.ws = (float*)workspace;
    auto make = [&](CUtensorMap* m, const void* basep, int C, int pitch, int halo_ext) {
        const uint64_t cB = (uint64_t)pitch * 2;
        const uint64_t bW = cB, bH = cB * d->W, bT = bH * d->H, bN = bT * d->T;
        uint64_t dims[5], str[4];
        dims[0] = C;
        dims[1] = d->W, str[0] = bW;
        if (p.spatial) {
            dims[2] = d->T, str[1] = bT;
            dims[3] = d->N, str[2] = bN;
            dims[4] = d->H, str[3] = bH;
        } else {
            dims[2] = d->H, str[1] = bH;
            dims[3] = d->N, str[2] = bN;
            dims[4] = d->T, str[3] = bT;
        }
        uint32_t box[5] = {64, (uint32_t)p.b[0], (uint32_t)p.b[1], (uint32_t)p.b[2], (uint32_t)halo_ext};
        return make_map(m, basep, 5, dims, str, box);
    };
    CUtensorMap mX, mDy;
    int rc = make(&mX, x, d->Cin, s.cinp, p.b[3] + 2);
    if (rc) return rc;
    rc = make(&mDy, dy, d->Cout, s.coutp, p.b[3]);
    if (rc) return rc;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    if (attr_err != cudaSuccess)
        return fail(ZSV_ERR_CUDA, "cudaFuncSetAttribute(wgrad halo) failed: %s", cudaGetErrorString(attr_err));
    zsv::launch(wgrad_halo_kernel, dim3(p.splits, p.ncopies), 192, p.smem, st, mX, mDy, a);
    ZSV_LAUNCH_CHECK("wgrad_halo_kernel");
    return ZSV_OK;
}
}  // namespace

namespace {
// split partials ws[split][tap][ci][co] -> dw[co][ci][tap] (fp32, the state-dict layout)
int launch_wgrad_finalize(const zsv_conv_desc* d, const Shape& s, const float* ws, float* dw, int splits, int ci_pitch,
                          int co_pitch, cudaStream_t st) {
    const long long wtotal = (long long)s.ftaps * ci_pitch * co_pitch;
    // Small tensors reduced over many splits (layers 1-2, stem): one thread per element keeps the whole GPU busy where the
    // tiled kernel would have a few dozen blocks each walking all splits (measured: 10 us vs 95 us on 144->64 with 147
    // splits); from ~1 M weights on (layers 3-4) the element-wise kernel's scattered 4-byte stores dominate and the
    // tiled one wins (44 -> 32 us on 512->1152, 28 -> 14 us on 1152->512).  Also the W-folded first layer
    // (tap = (dt,dh), ci = dw*8 + c).
    if (s.wfold || wtotal < (1LL << 20) || getenv("ZSV_DEBUG_FINALIZE_SMALL")) {
        const int blocks = (int)std::min<long long>(ceil_div_ll(wtotal, 256), 148 * 8);
        zsv::launch(wgrad_finalize_small_kernel, blocks, 256, 0, st, ws, dw, splits, s.ftaps, ci_pitch, co_pitch, d->Cin, d->Cout,
                                                            s.wfold ? d->kw : 0, s.ntaps);
        ZSV_LAUNCH_CHECK("wgrad_finalize_small_kernel");
        return ZSV_OK;
    }
    // input channels per block: runs of ~256-288 floats per output channel; fewer when that leaves SMs without a block
    // (small tensors summed over many splits)
    int ci_tile = std::max(8, std::min(64, (288 / s.ftaps) & ~7));
    while (ci_tile > 8 && (long long)ceil_div(d->Cout, 32) * ceil_div(d->Cin, ci_tile) < 2LL * sm_count()) ci_tile >>= 1;
    const size_t fsm = (size_t)32 * ((ci_tile * s.ftaps) | 1) * sizeof(float);
    if (fsm > 48 * 1024) return fail(ZSV_ERR_UNSUPPORTED, "wgrad finalize: %d taps do not fit the staging tile", s.ftaps);
    dim3 fgrid(ceil_div(d->Cout, 32), ceil_div(d->Cin, ci_tile));
    zsv::launch(wgrad_finalize_kernel, fgrid, 256, fsm, st, ws, dw, splits, s.ftaps, ci_pitch, co_pitch, d->Cin, d->Cout, ci_tile);
    ZSV_LAUNCH_CHECK("wgrad_finalize_kernel");
    return ZSV_OK;
}
}  // namespace

extern "C" size_t zsv_conv3d_wgrad_workspace(const zsv_conv_desc* d) {
    Shape s;
    if (check_desc(d, &s)) return 0;
    WgradPlan p;
    plan_wgrad(d, s, &p);
    const WgradHaloPlan hp = plan_wgrad_halo(d, s);
    const WgradPairPlan pp = plan_wgrad_pair(d, s);
    size_t need = p.ws_bytes;
    if (hp.ok) need = std::max(need, hp.ws_bytes);
    if (pp.ok) need = std::max(need, pp.ws_bytes);
    return need;
}

extern "C" int zsv_conv3d_wgrad(const zsv_conv_desc* d, const void* x, const void* dy, float* dw, void* workspace,
                                size_t workspace_bytes, void* stream) {
    Shape s;
    int rc = check_desc(d, &s);
    if (rc) return rc;
    if (!x || !dy || !dw || !workspace) return fail(ZSV_ERR_BAD_ARG, "wgrad: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const WgradHaloPlan hp = plan_wgrad_halo(d, s);
    if (hp.ok) {
        if (workspace_bytes < hp.ws_bytes)
            return fail(ZSV_ERR_WORKSPACE, "wgrad: workspace %zu < required %zu bytes", workspace_bytes, hp.ws_bytes);
        rc = launch_wgrad_halo(hp, d, s, x, dy, workspace, st);
        if (rc) return rc;
        return launch_wgrad_finalize(d, s, (const float*)workspace, dw, hp.splits, hp.ci_pitch, hp.co_pitch, st);
    }
    const WgradPairPlan pp = plan_wgrad_pair(d, s);
    if (pp.ok) {
        if (workspace_bytes < pp.ws_bytes)
            return fail(ZSV_ERR_WORKSPACE, "wgrad: workspace %zu < required %zu bytes", workspace_bytes, pp.ws_bytes);
        rc = launch_wgrad_pair(pp, d, s, x, dy, workspace, st);
        if (rc) return rc;
        return launch_wgrad_finalize(d, s, (const float*)workspace, dw, pp.splits, pp.ci_pitch, pp.co_pitch, st);
    }
    WgradPlan p;
    plan_wgrad(d, s, &p);
    if (workspace_bytes < p.ws_bytes)
        return fail(ZSV_ERR_WORKSPACE, "wgrad: workspace %zu < required %zu bytes", workspace_bytes, p.ws_bytes);

    WgradArgs a;
    memset(&a, 0, sizeof(a));
    const Box& b = p.box;
    a.bw = b.bw, a.bh = b.bh, a.bt = b.bt, a.bn = b.bn;
    a.tw = ceil_div(s.Wo, b.bw), a.th = ceil_div(s.Ho, b.bh), a.tt = ceil_div(s.To, b.bt), a.tn = ceil_div(d->N, b.bn);
    a.kchunks = p.kchunks;
    a.npanels = p.npanels;
    a.ntaps = s.ftaps;
    a.ci_store = p.ci_pitch;
    a.ci_pitch = p.ci_pitch;
    a.co_pitch = p.co_pitch;
    a.bn_tile = p.bn_tile;
    a.nbp = p.nbp;
    a.stages = p.stages;
    a.mt = p.mt;
    a.acc_stride = (p.bn_tile + 31) & ~31;
    a.tmem_cols = pow2_cols(p.mt * a.acc_stride);
    a.num_kb = p.num_kb;
    a.kb_per_split = p.kb_per_split;
    a.ws = (float*)workspace;

    CUtensorMap maps[kMaxMaps];
    int nmaps;
    rc = build_fwd_taps(d, s, x, b, a.taps, maps, &nmaps);
    if (rc) return rc;
    CUtensorMap mapB;
    rc = make_plain_map(&mapB, dy, d->N, s.To, s.Ho, s.Wo, d->Cout, s.coutp, b);
    if (rc) return rc;

    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(wgrad_mnmajor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    if (attr_err != cudaSuccess)
        return fail(ZSV_ERR_CUDA, "cudaFuncSetAttribute(wgrad) failed: %s", cudaGetErrorString(attr_err));
    const int smem = 1024 + p.stages * (2 * p.mt + p.nbp) * kPanelBytes + 16 * p.stages + 64;
    dim3 grid(p.m_tiles, p.n_tiles, p.splits);
    MapPack pack;
    for (int i = 0; i < kMaxMaps; ++i) pack.m[i] = maps[i];
    zsv::launch(wgrad_mnmajor_kernel, grid, 192, smem, st, pack, mapB, a);
    ZSV_LAUNCH_CHECK("wgrad_mnmajor_kernel");

    return launch_wgrad_finalize(d, s, (const float*)workspace, dw, p.splits, p.ci_pitch, p.co_pitch, st);
}
