// Microbenchmark: cycles per tcgen05.mma.cta_group::2 (CTA pair, M = 256) vs N, no-swizzle K-major
// operands, isolated issue loop (no TMA / epilogue traffic). Companion of tools/mma_bench.cu.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_bench_pair tools/mma_bench_pair.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../leela_b200/csrc/lb2_ptx.cuh"
using namespace lb2;

struct Cfg { int N, iters, two_acc; };

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) bench(Cfg c, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const bool leader = cluster_ctarank() == 0;
    for (int i = threadIdx.x; i < 150 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc_pair<512>(&tmem_slot);
    fence_proxy_async_smem();
    tc_fence_before_sync(); __syncthreads(); cluster_sync_all(); tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x < 32) {
        const uint32_t base = smem_u32(smem);
        const uint32_t a_base = base + 3 * 16, b_base = base + 96 * 1024;
        const int nh = c.N / 2;
        const uint32_t idesc = umma_idesc_f16(256, c.N);
        uint64_t ad[9], bd[9];
#pragma unroll
        for (int t = 0; t < 9; t++) {
            ad[t] = umma_desc_kmajor_noswizzle(a_base + t * 21 * 16, 304 * 16, 128);
            bd[t] = umma_desc_kmajor_noswizzle(b_base + t * nh * 32, nh * 16, 128);
        }
        const uint32_t d0 = tmem, d1 = tmem + (c.two_acc ? 128 : 0);
        long long t0 = 0, t1 = 0;
        if (leader) {
            if (elect_one()) {
#pragma unroll
                for (int t = 0; t < 9; t++) umma_f16<true>(d0, ad[t], bd[t], idesc, t > 0);
                umma_commit<true>(&bar);
            }
            __syncwarp();
        }
        mbar_wait(&bar, 0);
        t0 = clock64();
        if (leader) {
            if (elect_one()) {
                for (int i = 0; i < c.iters; i += 18) {
#pragma unroll
                    for (int t = 0; t < 9; t++) {
                        umma_f16<true>(d0, ad[t], bd[t], idesc, 1);
                        umma_f16<true>(d1, ad[t] + (c.two_acc ? 128 : 0), bd[t], idesc, 1);
                    }
                }
                umma_commit<true>(&bar);
            }
            __syncwarp();
        }
        mbar_wait(&bar, 1);
        t1 = clock64();
        if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before_sync(); __syncthreads(); cluster_sync_all();
    if (threadIdx.x < 32) tmem_dealloc_pair<512>(tmem);
}

int main() {
    long long* d_out; cudaMalloc(&d_out, 148 * sizeof(long long));
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    Cfg cfgs[] = {{64, 3600, 1}, {128, 3600, 1}, {256, 3600, 1}, {64, 3600, 0}, {128, 3600, 0}, {32, 3600, 1}, {96, 3600, 1}};
    printf("%5s %7s | %12s  ideal(per SM)\n", "N", "two_acc", "cyc/MMA(med)");
    for (auto& c : cfgs) {
        bench<<<148, 128, 160 * 1024>>>(c, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("config failed: %s\n", cudaGetErrorString(e)); return 1; }
        long long h[148]; cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
        for (int i = 0; i < 148; i++) for (int j = i + 1; j < 148; j++) if (h[j] < h[i]) { long long t = h[i]; h[i] = h[j]; h[j] = t; }
        printf("%5d %7d | %12.1f  %.0f\n", c.N, c.two_acc, (double)h[74] / c.iters, 128.0 * c.N / 256.0);
    }
    return 0;
}
