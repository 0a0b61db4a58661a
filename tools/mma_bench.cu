// Microbenchmark: cycles per tcgen05.mma (kind::f16, cta_group::1) as a function of M, N, the
// smem operand layout (no-swizzle K-major vs 128B-swizzle K-major), A-start alignment and the
// number of accumulators cycled. Operand contents are irrelevant (smem is zeroed).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_bench tools/mma_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../leela_b200/csrc/lb2_ptx.cuh"

using namespace lb2;

struct Cfg { int M, N, swz, a_shift16, n_acc, a_rows_apart, iters, b_shift16; };

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)(1) << 16;                  // LBO (unused for swizzled K-major)
    d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;  // SBO = 8 rows x 128 B
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)((addr >> 7) & 7) << 49;    // base offset when the start is not 1024-aligned
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
}

__global__ void __launch_bounds__(128, 1) bench(Cfg c, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < 180 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&tmem_slot);
    fence_proxy_async_smem();
    tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    long long t0 = 0, t1 = 0;
    if (threadIdx.x < 32) {
        const uint32_t base = smem_u32(smem);
        const uint32_t a_base = base + c.a_shift16 * 16;
        const uint32_t b_base = base + 96 * 1024 + c.b_shift16 * 16;
        const uint32_t idesc = (1u << 4) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
        // 9 operand positions like the taps of the real kernel, descriptors precomputed
        uint64_t ad[9], bd[9];
#pragma unroll
        for (int t = 0; t < 9; t++) {
            const uint32_t ao = t * (c.swz ? 1024u : (uint32_t)c.a_rows_apart * 16u);
            const uint32_t bo = t * (c.swz ? 2048u : (uint32_t)c.N * 32u);
            ad[t] = c.swz ? desc_sw128(a_base + ao) : umma_desc_kmajor_noswizzle(a_base + ao, 304 * 16, 128);
            bd[t] = c.swz ? desc_sw128(b_base + bo) : umma_desc_kmajor_noswizzle(b_base + bo, c.N * 16, 128);
        }
        const uint32_t d0 = tmem, d1 = tmem + (c.n_acc > 1 ? (c.N < 128 ? 128 : c.N) : 0);
        if (elect_one()) {
#pragma unroll
            for (int t = 0; t < 9; t++) umma_f16(d0, ad[t], bd[t], idesc, t > 0);
            umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        t0 = clock64();
        if (elect_one()) {
            for (int i = 0; i < c.iters; i += 18) {
#pragma unroll
                for (int t = 0; t < 9; t++) {
                    umma_f16(d0, ad[t], bd[t], idesc, 1);
                    umma_f16(d1, ad[t] + (c.n_acc > 1 ? 128 : 0), bd[t], idesc, 1);
                }
            }
            umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 1);
        t1 = clock64();
        if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before_sync(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

int main() {
    long long* d_out; cudaMalloc(&d_out, 148 * sizeof(long long));
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    Cfg cfgs[] = {
        // M,  N, swz, a_shift, n_acc, rows_apart, iters, b_shift
        {128, 64, 0, 0, 1, 8, 3600, 0},   {128, 128, 0, 0, 1, 8, 3600, 0},  {128, 256, 0, 0, 1, 8, 3600, 0},
        {128, 64, 0, 3, 1, 1, 3600, 0},   {128, 128, 0, 3, 1, 1, 3600, 0},  {128, 256, 0, 3, 1, 1, 3600, 0},
        {128, 128, 0, 3, 2, 1, 3600, 0},  {128, 128, 0, 3, 2, 20, 3600, 0},
        {128, 64, 1, 0, 1, 8, 3600, 0},   {128, 128, 1, 0, 1, 8, 3600, 0},  {128, 256, 1, 0, 1, 8, 3600, 0},
        {128, 128, 1, 8, 1, 8, 3600, 0},  {128, 256, 1, 8, 1, 8, 3600, 8},
        {64, 256, 0, 0, 1, 8, 3600, 0},   {64, 256, 0, 0, 1, 8, 3600, 3},   {128, 256, 0, 0, 1, 8, 3600, 3},
        {64, 128, 0, 0, 1, 8, 3600, 0},   {64, 256, 1, 0, 1, 8, 3600, 0},
    };
    printf("%5s %5s %4s %7s %5s %6s %7s | %10s %10s  ideal\n", "M", "N", "swz", "a_shift", "n_acc", "apart", "b_shift", "cyc/MMA(med)", "max");
    for (auto& c : cfgs) {
        bench<<<148, 128, 190 * 1024>>>(c, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("config failed: %s\n", cudaGetErrorString(e)); return 1; }
        long long h[148]; cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
        // median and max over CTAs
        for (int i = 0; i < 148; i++) for (int j = i + 1; j < 148; j++) if (h[j] < h[i]) { long long t = h[i]; h[i] = h[j]; h[j] = t; }
        double ideal = (double)(c.M < 128 ? 128 : c.M) * c.N / 256.0;
        printf("%5d %5d %4d %7d %5d %6d %7d | %10.1f %10.1f  %.0f\n", c.M, c.N, c.swz, c.a_shift16, c.n_acc, c.a_rows_apart, c.b_shift16,
               (double)h[74] / c.iters, (double)h[147] / c.iters, ideal);
    }
    return 0;
}
