// Microbenchmark: how fast does ONE thread feed tcgen05.mma (kind::f16, bf16, K = 16) on sm_100a?
// cycles per MMA for cta_group::1 (M = 128) and cta_group::2 (M = 256) over N, with K-major SWIZZLE_128B operands in
// shared memory, issued back to back by one thread (4 k-steps per 128-byte row like the convolution kernels), timed with
// clock64 from the first issue to the completion of the last (tcgen05.commit -> mbarrier).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/mma_rate tools/mma_rate.cu
// run:   tools/bin/mma_rate            (prints a table; profiles/r02_mma_rate.txt is one such run)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../zeroshotvideoclassification_b200/csrc/zsv_ptx.cuh"
using namespace zsv;

// a_sbo: byte distance between 8-row groups of A (1024 = dense atoms; 1280 = the halo kernel's widened W box);
// a_row0: start row of A inside its tile (W tap offset); spread: distinct A tiles cycled through (ring stages)
template <bool k2>
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int iters, uint32_t a_sbo, uint32_t a_row0, int spread,
                                                      int gap_instrs, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint32_t tslot;
    __shared__ alignas(8) unsigned long long bar_storage;
    const uint32_t bar = smem_u32(&bar_storage);
    const int warp = threadIdx.x >> 5;
    // operands: zeros (finite values; timing does not depend on them)
    for (uint32_t i = threadIdx.x; i < (200u * 1024u) / 16u; i += blockDim.x)
        reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)))[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        if (k2) tmem_alloc2(smem_u32(&tslot), 512);
        else tmem_alloc(smem_u32(&tslot), 512);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    if (k2) cluster_sync_all();
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tslot;
    const uint32_t rank = k2 ? cluster_ctarank() : 0u;
    long long dt = 0;
    if (warp == 1 && rank == 0) {
        const uint32_t leader = elect_one();
        if (leader) {
            const uint32_t idesc = umma_idesc_bf16(k2 ? 256 : 128, N, 0, 0);
            const uint32_t a_hi = umma_desc_hi(a_sbo, 2), b_hi = umma_desc_hi(1024, 2);
            const uint32_t a_tile = 128u * 160u;   // room for 128 rows at SBO 1280
            const uint32_t b0 = umma_desc_lo(base + 4u * a_tile);
            const long long t0 = clock64();
            int s = 0;
            uint32_t filler = 0;
            for (int i = 0; i < iters; i += 4) {
                const uint32_t a0 = umma_desc_lo(base + s * a_tile + a_row0 * 128u);
                if (++s >= spread) s = 0;
                if (k2) {
                    umma2_bf16_lohi(tmem, a0, a_hi, b0, b_hi, idesc, 1u);
                    umma2_bf16_lohi(tmem, a0 + 2u, a_hi, b0 + 2u, b_hi, idesc, 1u);
                    umma2_bf16_lohi(tmem, a0 + 4u, a_hi, b0 + 4u, b_hi, idesc, 1u);
                    umma2_bf16_lohi(tmem, a0 + 6u, a_hi, b0 + 6u, b_hi, idesc, 1u);
                } else {
                    umma_bf16_lohi(tmem, a0, a_hi, b0, b_hi, idesc, 1u);
                    umma_bf16_lohi(tmem, a0 + 2u, a_hi, b0 + 2u, b_hi, idesc, 1u);
                    umma_bf16_lohi(tmem, a0 + 4u, a_hi, b0 + 4u, b_hi, idesc, 1u);
                    umma_bf16_lohi(tmem, a0 + 6u, a_hi, b0 + 6u, b_hi, idesc, 1u);
                }
                // optional scalar filler between groups (models loop overhead of the real issue loops)
                for (int g = 0; g < gap_instrs; ++g) asm volatile("add.u32 %0, %0, 1;" : "+r"(filler));
            }
            if (k2) umma2_commit_mc(bar, 1);
            else umma_commit(bar);
            mbar_wait(bar, 0);
            dt = clock64() - t0 + (filler == 0xffffffffu);
        }
        __syncwarp();
    }
    tc_fence_before();
    if (k2) cluster_sync_all();
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (k2) tmem_dealloc2(tmem, 512);
        else tmem_dealloc(tmem, 512);
    }
    if (threadIdx.x == 32 && rank == 0) out[blockIdx.x] = dt;
}

template <bool k2>
double run(int grid, int N, int iters, uint32_t a_sbo, uint32_t a_row0, int spread, int gap, long long* dout) {
    const int smem = 210 * 1024;
    cudaFuncSetAttribute(rate_kernel<k2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid), cfg.blockDim = dim3(128), cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = k2 ? 2 : 1, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    cudaMemset(dout, 0, sizeof(long long) * grid);
    for (int rep = 0; rep < 2; ++rep) {
        cudaError_t e = cudaLaunchKernelEx(&cfg, rate_kernel<k2>, N, iters, a_sbo, a_row0, spread, gap, dout);
        if (e != cudaSuccess) { printf("launch: %s\n", cudaGetErrorString(e)); exit(1); }
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("sync: %s\n", cudaGetErrorString(e)); exit(1); }
    }
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), dout, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double sum = 0;
    int n = 0;
    for (int i = 0; i < grid; ++i)
        if (h[i] > 0) sum += (double)h[i], ++n;
    return sum / n / iters;
}

int main() {
    long long* dout;
    cudaMalloc(&dout, sizeof(long long) * 1024);
    const int iters = 4096;
    const int Ns[] = {16, 32, 48, 64, 96, 128, 144, 192, 240, 256};
    printf("cycles per tcgen05.mma (bf16, K=16), one issuing thread, %d MMAs back to back; math = M*N*16/8192 clk per SM at 8192 MAC/clk\n", iters);
    printf("%-44s", "variant \\ N");
    for (int N : Ns) printf("%7d", N);
    printf("\n");
    struct V { const char* name; bool k2; int grid; uint32_t sbo, row0; int spread, gap; };
    const V vs[] = {
        {"cta_group::1 M=128, 148 CTAs", false, 148, 1024, 0, 1, 0},
        {"cta_group::1 M=128, 1 CTA", false, 1, 1024, 0, 1, 0},
        {"cta_group::2 M=256, 74 pairs", true, 148, 1024, 0, 1, 0},
        {"cta_group::2 M=256, 1 pair", true, 2, 1024, 0, 1, 0},
        {"cta_group::1, A over 4 tiles", false, 148, 1024, 0, 4, 0},
        {"cta_group::2, A over 4 tiles", true, 148, 1024, 0, 4, 0},
        {"cta_group::1, A SBO 1280 start row 1", false, 148, 1280, 1, 1, 0},
        {"cta_group::2, A SBO 1280 start row 1", true, 148, 1280, 1, 1, 0},
        {"cta_group::2, A SBO 1280 row 1, 4 tiles", true, 148, 1280, 1, 4, 0},
        {"cta_group::1, +8 filler instr / 4 MMA", false, 148, 1024, 0, 1, 8},
        {"cta_group::2, +8 filler instr / 4 MMA", true, 148, 1024, 0, 1, 8},
        {"cta_group::2, +24 filler instr / 4 MMA", true, 148, 1024, 0, 1, 24},
    };
    for (const V& v : vs) {
        printf("%-44s", v.name);
        for (int N : Ns) {
            if (v.k2 && (N % 16)) { printf("%7s", "-"); continue; }
            double c = v.k2 ? run<true>(v.grid, N, iters, v.sbo, v.row0, v.spread, v.gap, dout)
                            : run<false>(v.grid, N, iters, v.sbo, v.row0, v.spread, v.gap, dout);
            printf("%7.1f", c);
        }
        printf("\n");
    }
    return 0;
}
