// Hardware check for the fp16 + fp8-correction scheme ("lite" split operands):
//   1. a kind::f16 MMA (K = 16) and a kind::f8f6f4 MMA (e4m3, K = 32) may accumulate into the SAME fp32 TMEM
//      accumulator, both reading K-major no-swizzle core-matrix operands of identical byte geometry
//      (8 rows x 16 bytes per core matrix: 8 fp16 or 16 fp8 channels);
//   2. the issue rate of alternating kinds against one kind alone;
//   3. HOW PRECISELY kind::f8f6f4 adds into an accumulator that already holds a large value: the e4m3 products of one
//      instruction are summed and aligned to the accumulator's exponent with a limited number of bits — an addend
//      2^-s of the accumulator survives only for small s (printed below). kind::f16 adds the same addend exactly. This
//      is why the lite mode keeps its e4m3 correction terms in an accumulator of their own.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_mixed_test tools/mma_mixed_test.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include "../leela_b200/csrc/lb2_ptx.cuh"

using namespace lb2;

__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

constexpr int M = 128, N = 64;
// smem: A16 [2 chunks][M rows][16 B] | B16 [2][N][16 B] | A8 [2][M][16 B] | B8 [2][N][16 B]
constexpr int kA = 2 * M * 16, kB = 2 * N * 16;

__global__ void __launch_bounds__(128, 1) mixed(const uint8_t* a16, const uint8_t* b16, const uint8_t* a8, const uint8_t* b8,
                                                float* out, int mode, int iters, long long* cycles) {
    __shared__ __align__(1024) uint8_t smem[2 * kA + 2 * kB];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < kA; i += blockDim.x) { smem[i] = a16[i]; smem[kA + kB + i] = a8[i]; }
    for (int i = threadIdx.x; i < kB; i += blockDim.x) { smem[kA + i] = b16[i]; smem[2 * kA + kB + i] = b8[i]; }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc<64>(&tmem_slot);
    fence_proxy_async_smem();
    tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t base = smem_u32(smem);
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint64_t da16 = umma_desc_kmajor_noswizzle(base, M * 16, 128), db16 = umma_desc_kmajor_noswizzle(base + kA, N * 16, 128);
    const uint64_t da8 = umma_desc_kmajor_noswizzle(base + kA + kB, M * 16, 128), db8 = umma_desc_kmajor_noswizzle(base + 2 * kA + kB, N * 16, 128);
    if (threadIdx.x < 32) {
        long long t0 = clock64();
        if (elect_one()) {
            if (mode == 0 || mode == 4) {   // correctness: one of each into the same accumulator
                umma_f16(tmem, da16, db16, idesc, 0);
                umma_f8(tmem, da8, db8, idesc, 1);
            } else if (mode == 5) {         // the same small addend through kind::f16 (operands in the fp16 buffers' second half)
                umma_f16(tmem, da16, db16, idesc, 0);
                umma_f16(tmem, da8, db8, idesc, 1);
            } else {
                for (int i = 0; i < iters; i += 2) {
                    if (mode == 1) { umma_f16(tmem, da16, db16, idesc, 1); umma_f16(tmem, da16, db16, idesc, 1); }
                    if (mode == 2) { umma_f8(tmem, da8, db8, idesc, 1); umma_f8(tmem, da8, db8, idesc, 1); }
                    if (mode == 3) { umma_f16(tmem, da16, db16, idesc, 1); umma_f8(tmem, da8, db8, idesc, 1); }
                }
            }
            umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        if (threadIdx.x == 0) cycles[0] = clock64() - t0;
    }
    tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
    if (mode == 0 || mode == 4 || mode == 5) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int c = 0; c < N; c += 8) {
            uint32_t r[8];
            tmem_ld_32x8(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
            tmem_ld_wait();
            for (int e = 0; e < 8; e++) out[(warp * 32 + lane) * N + c + e] = __uint_as_float(r[e]);
        }
    }
    tc_fence_before_sync(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<64>(tmem);
}

static float e4m3_to_float(uint8_t v) {
    __half_raw h = __nv_cvt_fp8_to_halfraw(v, __NV_E4M3);
    return __half2float(*reinterpret_cast<__half*>(&h));
}

int main() {
    std::vector<uint8_t> a16(kA), b16(kB), a8(kA), b8(kB);
    std::vector<float> fa16(M * 16), fb16(N * 16), fa8(M * 32), fb8(N * 32);
    srand(7);
    auto rnd = []() { return (float)((rand() % 33) - 16) / 8.0f; };   // multiples of 1/8 in [-2, 2]: exact in fp16 and e4m3
    for (int r = 0; r < M; r++) for (int k = 0; k < 16; k++) {
        float v = rnd(); fa16[r * 16 + k] = v;
        reinterpret_cast<__half*>(a16.data())[((k / 8) * M + r) * 8 + k % 8] = __float2half(v);
    }
    for (int r = 0; r < N; r++) for (int k = 0; k < 16; k++) {
        float v = rnd(); fb16[r * 16 + k] = v;
        reinterpret_cast<__half*>(b16.data())[((k / 8) * N + r) * 8 + k % 8] = __float2half(v);
    }
    for (int r = 0; r < M; r++) for (int k = 0; k < 32; k++) {
        float v = rnd(); uint8_t q = __nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3); fa8[r * 32 + k] = e4m3_to_float(q);
        a8[((k / 16) * M + r) * 16 + k % 16] = q;
    }
    for (int r = 0; r < N; r++) for (int k = 0; k < 32; k++) {
        float v = rnd(); uint8_t q = __nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3); fb8[r * 32 + k] = e4m3_to_float(q);
        b8[((k / 16) * N + r) * 16 + k % 16] = q;
    }
    uint8_t *d_a16, *d_b16, *d_a8, *d_b8; float* d_out; long long* d_cyc;
    cudaMalloc(&d_a16, kA); cudaMalloc(&d_b16, kB); cudaMalloc(&d_a8, kA); cudaMalloc(&d_b8, kB);
    cudaMalloc(&d_out, M * N * sizeof(float)); cudaMalloc(&d_cyc, 8);
    cudaMemcpy(d_a16, a16.data(), kA, cudaMemcpyHostToDevice); cudaMemcpy(d_b16, b16.data(), kB, cudaMemcpyHostToDevice);
    cudaMemcpy(d_a8, a8.data(), kA, cudaMemcpyHostToDevice); cudaMemcpy(d_b8, b8.data(), kB, cudaMemcpyHostToDevice);
    mixed<<<1, 128>>>(d_a16, d_b16, d_a8, d_b8, d_out, 0, 0, d_cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> out(M * N);
    cudaMemcpy(out.data(), d_out, out.size() * sizeof(float), cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int r = 0; r < M; r++) for (int c = 0; c < N; c++) {
        double want = 0;
        for (int k = 0; k < 16; k++) want += (double)fa16[r * 16 + k] * fb16[c * 16 + k];
        for (int k = 0; k < 32; k++) want += (double)fa8[r * 32 + k] * fb8[c * 32 + k];
        worst = fmax(worst, fabs(want - out[r * N + c]));
    }
    printf("mixed f16 (K=16) + e4m3 (K=32) into one accumulator: max |err| = %g  -> %s\n", worst, worst == 0 ? "EXACT" : "MISMATCH");
    // 3. big accumulator (fp16 MMA: row r, column c holds 2^big) + one small e4m3 product 2^-s relative to it
    for (int via_f16 = 0; via_f16 < 2; via_f16++) {
        printf("small addend into an accumulator of 2^17 through %s: survives (exactly) down to 2^-", via_f16 ? "kind::f16    " : "kind::f8f6f4 ");
        int last_ok = -1;
        for (int sh = 4; sh <= 24; sh++) {
            std::vector<uint8_t> A16(kA, 0), B16(kB, 0), A8(kA, 0), B8(kB, 0);
            // fp16: A[r][0] = 2^9, B[c][0] = 2^8 -> 2^17 everywhere
            for (int r = 0; r < M; r++) reinterpret_cast<__half*>(A16.data())[r * 8] = __float2half(512.0f);
            for (int c = 0; c < N; c++) reinterpret_cast<__half*>(B16.data())[c * 8] = __float2half(256.0f);
            // small term: a = 2^(9 - sh/2 ...) split over the two operands so that both stay inside e4m3's normal range
            const int ea = 8 - (sh + 1) / 2, eb = 9 - sh / 2;   // a * b = 2^(17 - sh)
            if (via_f16) {
                for (int r = 0; r < M; r++) reinterpret_cast<__half*>(A8.data())[r * 8] = __float2half(ldexpf(1.0f, ea));
                for (int c = 0; c < N; c++) reinterpret_cast<__half*>(B8.data())[c * 8] = __float2half(ldexpf(1.0f, eb));
            } else {
                if (ea < -6 || eb < -6) break;
                for (int r = 0; r < M; r++) A8[r * 16] = __nv_cvt_float_to_fp8(ldexpf(1.0f, ea), __NV_SATFINITE, __NV_E4M3);
                for (int c = 0; c < N; c++) B8[c * 16] = __nv_cvt_float_to_fp8(ldexpf(1.0f, eb), __NV_SATFINITE, __NV_E4M3);
            }
            cudaMemcpy(d_a16, A16.data(), kA, cudaMemcpyHostToDevice); cudaMemcpy(d_b16, B16.data(), kB, cudaMemcpyHostToDevice);
            cudaMemcpy(d_a8, A8.data(), kA, cudaMemcpyHostToDevice); cudaMemcpy(d_b8, B8.data(), kB, cudaMemcpyHostToDevice);
            mixed<<<1, 128>>>(d_a16, d_b16, d_a8, d_b8, d_out, via_f16 ? 5 : 4, 0, d_cyc);
            if ((e = cudaDeviceSynchronize()) != cudaSuccess) { printf("failed: %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(out.data(), d_out, out.size() * sizeof(float), cudaMemcpyDeviceToHost);
            const float want = 131072.0f + ldexpf(1.0f, 17 - sh);
            if (out[0] == want && out[M * N - 1] == want) last_ok = sh; else { printf("%d (at 2^-%d the accumulator reads %.6f instead of %.6f)\n", last_ok, sh, out[0], want); break; }
            if (sh == 24) printf("%d (all tested)\n", last_ok);
        }
    }
    const char* names[] = {"", "f16 only", "f8 only", "alternating f16/f8"};
    for (int mode = 1; mode <= 3; mode++) {
        const int iters = 4000;
        mixed<<<1, 128>>>(d_a16, d_b16, d_a8, d_b8, d_out, mode, iters, d_cyc);
        if ((e = cudaDeviceSynchronize()) != cudaSuccess) { printf("mode %d failed: %s\n", mode, cudaGetErrorString(e)); return 1; }
        long long cyc; cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-22s M=%d N=%d: %.1f cycles per MMA\n", names[mode], M, N, (double)cyc / iters);
    }
    return worst == 0 ? 0 : 2;
}
