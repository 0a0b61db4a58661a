// lb2_kernels.cu — sm_100a kernels of the policy/value evaluator.
//
//   expand_planes_kernel  bit-planes + symmetry -> fp16 input of the first conv
//                         (Network.cpp:765-773 plane expansion + rotate_nn_idx)
//   trunk_kernel          every conv layer with c_out > 1 of both nets: shifted-row implicit
//                         GEMM on tcgen05 tensor cores, accumulators in TMEM, operands staged
//                         by TMA, bias + ELU fused in the epilogue (replaces im2col +
//                         cblas_sgemm + ELU sweep, Network.cpp:345-393, and the OpenCL
//                         convolve5/convolve3/merge kernels, OpenCL.cpp:25-377)
//   policy_head_kernel    last conv (C -> 1) + ELU + softmax(T) + un-rotation
//                         (Network.cpp:806-808, 450-469, 820-823)
//   value_head_kernel     last conv (C -> 1) + ELU + ip 361->H + ELU + ip H->1 + (1+tanh)/2
//                         (Network.cpp:731-737, 395-423; OpenCL innerproduct OpenCL.cpp:407-438)
#include "lb2_kernels.cuh"
#include "lb2_ptx.cuh"

namespace lb2 {

// Network::rotate_nn_idx (Network.cpp:1348-1379): bit2 swaps x/y first, bit0 flips y, bit1 flips x.
__device__ __forceinline__ int rotate_idx(int v, int s) {
    int x = v % kBoard, y = v / kBoard;
    if (s & 4) { int t = x; x = y; y = t; }
    if (s & 1) y = kBoard - 1 - y;
    if (s & 2) x = kBoard - 1 - x;
    return y * kBoard + x;
}
// Network::rev_rotate_nn_idx (Network.cpp:1341-1346)
__device__ __forceinline__ int rev_rotate_idx(int v, int s) {
    const int inv = (s == 5) ? 6 : (s == 6) ? 5 : s;
    return rotate_idx(v, inv);
}

__device__ __forceinline__ float elu1(float v) { return v > 0.0f ? v : (__expf(v) - 1.0f); }

// ------------------------------------------------------------------------------------------
// expand: one thread per row of the S=21 row space; writes 32 channels = 4 chunks of 8 fp16.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) expand_planes_kernel(const uint32_t* __restrict__ planes,
                                                            const uint8_t* __restrict__ rotation, int n,
                                                            __half* __restrict__ x0, int chunk_rows) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n * 441) return;
    const int pos = row / 441, rem = row - pos * 441;
    const int y = rem / 21, x = rem - y * 21;
    uint32_t bits = 0;
    if (x < kBoard && y < kBoard) {
        const int src = rotate_idx(y * kBoard + x, rotation[pos] & 7);
        bits = planes[(size_t)pos * kPoints + src];
    }
    const uint32_t one = 0x3C00u;  // fp16 1.0
#pragma unroll
    for (int c8 = 0; c8 < 4; c8++) {
        uint32_t b = bits >> (8 * c8);
        uint4 v;
        v.x = ((b & 1u) ? one : 0u) | ((b & 2u) ? (one << 16) : 0u);
        v.y = ((b & 4u) ? one : 0u) | ((b & 8u) ? (one << 16) : 0u);
        v.z = ((b & 16u) ? one : 0u) | ((b & 32u) ? (one << 16) : 0u);
        v.w = ((b & 64u) ? one : 0u) | ((b & 128u) ? (one << 16) : 0u);
        *reinterpret_cast<uint4*>(x0 + ((size_t)c8 * chunk_rows + row) * 8) = v;
    }
}

// ------------------------------------------------------------------------------------------
// trunk: persistent, warp-specialised. Work item = (layer job, 256-row tile).
//   K loop = for each 16-channel slab, for each tap group: one smem stage holds the A slab
//   ([2 chunks][256+2*halo rows][8] fp16, one TMA box) and the B block ([tap][2][N][8] fp16,
//   one bulk copy). For every tap the MMA reads the SAME A slab at a row offset
//   dy*S+dx — the im2col matrix is never materialised and each activation byte is fetched
//   from L2 once per item instead of k*k times.
// ------------------------------------------------------------------------------------------
struct TapGroups {
    int n;
    int begin[3], end[3];
};
__device__ __forceinline__ TapGroups tap_groups(int ksize) {
    TapGroups g;
    if (ksize == 3) { g.n = 1; g.begin[0] = 0; g.end[0] = 9; g.begin[1] = g.end[1] = g.begin[2] = g.end[2] = 0; }
    else { g.n = 3; g.begin[0] = 0; g.end[0] = 9; g.begin[1] = 9; g.end[1] = 17; g.begin[2] = 17; g.end[2] = 25; }
    return g;
}

__device__ __forceinline__ void wait_dependencies(const LayerJob& J, const LayerJob* jobs, int tile, uint32_t epoch) {
    const int r_lo = tile * kTileRows - J.halo;
    const int r_hi = tile * kTileRows + kTileRows - 1 + J.halo;
    int lo, hi;
    if (!J.dep_remap) {
        lo = r_lo < 0 ? 0 : r_lo / kTileRows;
        hi = r_hi / kTileRows;
    } else {  // producer enumerates the S=21 space (441 rows / position), we the S=20 space (400)
        int p_lo = (r_lo < 0 ? 0 : r_lo) / 400;
        int p_hi = r_hi / 400;
        if (p_hi > J.n_pos - 1) p_hi = J.n_pos - 1;
        if (p_lo > p_hi) p_lo = p_hi;
        lo = (p_lo * 441) / kTileRows;
        hi = ((p_hi + 1) * 441 - 1) / kTileRows;
    }
    if (hi > J.dep_n_items - 1) hi = J.dep_n_items - 1;
    const uint32_t* flags = jobs[J.dep_job].flags;
    for (int i = lo; i <= hi; i++) {
        while (ld_acquire_gpu(flags + i) != epoch) __nanosleep(64);
    }
    fence_proxy_async();  // order the TMA (async proxy) reads after the acquire
}

__global__ void __launch_bounds__(kTrunkThreads, 1) trunk_kernel(const __grid_constant__ TrunkParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tfull_bar = empty_bar + kStages;   // [2] accumulator ready   (MMA -> epilogue)
    uint64_t* tempty_bar = tfull_bar + 2;        // [2] accumulator drained (epilogue -> MMA)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 127u) __trap();  // TMA destinations need 128-byte alignment
        for (int i = 0; i < kStages; i++) { mbar_init(full_bar + i, 1); mbar_init(empty_bar + i, 1); }
        for (int i = 0; i < 2; i++) { mbar_init(tfull_bar + i, 1); mbar_init(tempty_bar + i, 4); }
        fence_mbar_init();
        fence_proxy_async_smem();
        for (int i = 0; i < kMaxTensorMaps; i++) tma_prefetch_desc(&P.tmaps[i]);
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const LayerJob* __restrict__ jobs = P.jobs;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0; int j = 0;
            for (int q = P.item_begin + blockIdx.x; q < P.item_end; q += gridDim.x) {
                while (q >= jobs[j].item_base + jobs[j].n_items) j++;
                const LayerJob J = jobs[j];
                const int tile = q - J.item_base;
                if (P.use_flags && J.dep_job >= 0) wait_dependencies(J, jobs, tile, P.epoch);
                const int rows_halo = kTileRows + 2 * J.halo;
                const uint32_t a_bytes = rows_halo * 32;
                const int row0_8 = (tile * kTileRows - J.halo) / 8;  // exact: both multiples of 8
                const TapGroups G = tap_groups(J.ksize);
                const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(J.wpk);
                for (int s = 0; s < J.n_slabs; s++) {
                    for (int g = 0; g < G.n; g++) {
                        const uint32_t b_bytes = (G.end[g] - G.begin[g]) * J.n_out * 32;
                        mbar_wait(empty_bar + stage, phase ^ 1);
                        uint8_t* sa = smem + stage * kStageBytes;
                        mbar_arrive_expect_tx(full_bar + stage, a_bytes + b_bytes);
                        tma_load_3d(sa, &P.tmaps[J.tmap], full_bar + stage, 0, row0_8, 2 * s);
                        bulk_load_1d(sa + kASlabBytes, wsrc, b_bytes, full_bar + stage);
                        wsrc += b_bytes;
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0; int j = 0; uint32_t it = 0;
            for (int q = P.item_begin + blockIdx.x; q < P.item_end; q += gridDim.x, it++) {
                while (q >= jobs[j].item_base + jobs[j].n_items) j++;
                const LayerJob J = jobs[j];
                const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
                mbar_wait(tempty_bar + acc, acc_phase ^ 1);
                tc_fence_after_sync();
                const uint32_t idesc = umma_idesc_f16(128, J.n_out);
                const int rows_halo = kTileRows + 2 * J.halo;
                uint32_t lbo_a = rows_halo * 16, lbo_b = J.n_out * 16, sbo = 128;
                if (P.debug_flags & 1) { sbo = lbo_a; lbo_a = 128; }
                const int pad = J.ksize >> 1;
                const TapGroups G = tap_groups(J.ksize);
                const uint32_t d0 = tmem_base + (acc * 2 + 0) * 128;
                const uint32_t d1 = tmem_base + (acc * 2 + 1) * 128;
                bool first = true;
                for (int s = 0; s < J.n_slabs; s++) {
                    for (int g = 0; g < G.n; g++) {
                        mbar_wait(full_bar + stage, phase);
                        tc_fence_after_sync();
                        const uint32_t a_addr = smem_u32(smem + stage * kStageBytes);
                        const uint32_t b_addr = a_addr + kASlabBytes;
                        for (int t = G.begin[g]; t < G.end[g]; t++) {
                            const int kr = t / J.ksize, kc = t - kr * J.ksize;
                            const int off = (kr - pad) * J.S + (kc - pad);
                            const uint64_t bdesc =
                                (P.debug_flags & 1)
                                    ? umma_desc_kmajor_noswizzle(b_addr + (t - G.begin[g]) * J.n_out * 32, 128, lbo_b)
                                    : umma_desc_kmajor_noswizzle(b_addr + (t - G.begin[g]) * J.n_out * 32, lbo_b, 128);
                            const uint32_t a0 = a_addr + (J.halo + off) * 16;
                            umma_f16(d0, umma_desc_kmajor_noswizzle(a0, lbo_a, sbo), bdesc, idesc, first ? 0u : 1u);
                            umma_f16(d1, umma_desc_kmajor_noswizzle(a0 + 128 * 16, lbo_a, sbo), bdesc, idesc,
                                     first ? 0u : 1u);
                            first = false;
                        }
                        umma_commit(empty_bar + stage);  // frees the smem stage when these MMAs finish
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
                umma_commit(tfull_bar + acc);  // accumulator complete -> epilogue
            }
        }
    } else {
        // ================================ epilogue ====================================
        const int quad = warp & 3;  // TMEM lane quadrant this warp may read
        int j = 0; uint32_t it = 0;
        for (int q = P.item_begin + blockIdx.x; q < P.item_end; q += gridDim.x, it++) {
            while (q >= jobs[j].item_base + jobs[j].n_items) j++;
            const LayerJob& J = jobs[j];
            const int tile = q - J.item_base;
            const int S = J.S, SS = S * S, n_out = J.n_out, chunk_rows = J.out_chunk_rows;
            const float* __restrict__ bias = J.bias;
            __half* __restrict__ out = J.out;
            const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
            mbar_wait(tfull_bar + acc, acc_phase);
            tc_fence_after_sync();
#pragma unroll 1
            for (int h = 0; h < 2; h++) {
                const int row = tile * kTileRows + h * 128 + quad * 32 + lane;
                const int pos = row / SS, rem = row - pos * SS;
                const int y = rem / S, x = rem - y * S;
                const bool valid = (x < kBoard) && (y < kBoard);
                int out_row = row;
                bool store = true;
                if (J.remap) { out_row = pos * 400 + y * 20 + x; store = valid && pos < J.n_pos; }
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + (acc * 2 + h) * 128;
#pragma unroll 1
                for (int cc = 0; cc < n_out / 32; cc++) {
                    uint32_t r[32];
                    tmem_ld_32x32(taddr + cc * 32, r);
                    tmem_ld_wait();
                    uint32_t pk[16];
#pragma unroll
                    for (int e = 0; e < 16; e++) {
                        float v0 = __uint_as_float(r[2 * e]) + __ldg(bias + cc * 32 + 2 * e);
                        float v1 = __uint_as_float(r[2 * e + 1]) + __ldg(bias + cc * 32 + 2 * e + 1);
                        v0 = valid ? elu1(v0) : 0.0f;
                        v1 = valid ? elu1(v1) : 0.0f;
                        __half2 hh = __floats2half2_rn(v0, v1);
                        pk[e] = *reinterpret_cast<uint32_t*>(&hh);
                    }
                    if (store) {
#pragma unroll
                        for (int c8 = 0; c8 < 4; c8++) {
                            uint4 v = make_uint4(pk[4 * c8], pk[4 * c8 + 1], pk[4 * c8 + 2], pk[4 * c8 + 3]);
                            *reinterpret_cast<uint4*>(out + ((size_t)(cc * 4 + c8) * chunk_rows + out_row) * 8) = v;
                        }
                    }
                }
            }
            // accumulator drained: hand it back to the MMA warp
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar + acc);
            if (P.use_flags) {
                __threadfence();
                named_bar_sync(1, 128);
                if (threadIdx.x == 64) {
                    fence_proxy_async();  // generic-proxy stores -> visible to other CTAs' TMA loads
                    st_release_gpu(J.flags + tile, P.epoch);
                }
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------
// policy head: one CTA per position. conv C->1 (3x3, fp32 weights) + bias + ELU, softmax with
// temperature over the 361 points, un-rotate.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float head_conv_pixel(const __half* __restrict__ act, int chunk_rows, int c_in,
                                                 const float* __restrict__ w_s /*[9][c_in]*/, int base_row, int y,
                                                 int x) {
    float acc = 0.0f;
#pragma unroll 1
    for (int t = 0; t < 9; t++) {
        const int dy = t / 3 - 1, dx = t % 3 - 1;
        const int row = base_row + (y + dy) * 20 + (x + dx);
        if (row < 0) continue;  // above the first position: implicit zero padding
        const float* wt = w_s + t * c_in;
        for (int c8 = 0; c8 < c_in / 8; c8++) {
            const uint4 v = *reinterpret_cast<const uint4*>(act + ((size_t)c8 * chunk_rows + row) * 8);
            const __half2* hp = reinterpret_cast<const __half2*>(&v);
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const float2 f = __half22float2(hp[e]);
                acc = fmaf(f.x, wt[c8 * 8 + 2 * e], acc);
                acc = fmaf(f.y, wt[c8 * 8 + 2 * e + 1], acc);
            }
        }
    }
    return acc;
}

__device__ __forceinline__ float warp_max(float v) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(384) policy_head_kernel(const __half* __restrict__ act, int chunk_rows, int c_in,
                                                          const float* __restrict__ w /*[c_in][3][3]*/,
                                                          const float* __restrict__ bias,
                                                          const uint8_t* __restrict__ rotation, float temp,
                                                          float* __restrict__ probs) {
    extern __shared__ float hs[];
    float* w_s = hs;                  // [9][c_in]
    float* sm = hs + 9 * c_in;        // [361]
    __shared__ float red[12];
    const int pos = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < 9 * c_in; i += blockDim.x) {
        const int t = i / c_in, c = i - t * c_in;
        w_s[i] = w[c * 9 + t];
    }
    __syncthreads();
    float logit = -INFINITY;
    if (tid < kPoints) {
        const int y = tid / kBoard, x = tid - y * kBoard;
        logit = elu1(bias[0] + head_conv_pixel(act, chunk_rows, c_in, w_s, pos * 400, y, x));
    }
    // softmax(x / T): p = exp(x/T - max/T) / sum  (Network.cpp:450-469)
    float m = warp_max(logit);
    if ((tid & 31) == 0) red[tid >> 5] = m;
    __syncthreads();
    m = red[0];
    for (int i = 1; i < 12; i++) m = fmaxf(m, red[i]);
    __syncthreads();
    const float e = (tid < kPoints) ? expf(logit / temp - m / temp) : 0.0f;
    float s = warp_sum(e);
    if ((tid & 31) == 0) red[tid >> 5] = s;
    __syncthreads();
    s = 0.0f;
    for (int i = 0; i < 12; i++) s += red[i];
    if (tid < kPoints) sm[tid] = e / s;
    __syncthreads();
    if (tid < kPoints) probs[(size_t)pos * kPoints + tid] = sm[rev_rotate_idx(tid, rotation[pos] & 7)];
}

// ------------------------------------------------------------------------------------------
// value head: kValueGroup positions per CTA so the 361xH inner-product matrix is read once per
// group. conv C->1 + ELU -> v[361]; h = ELU(W1 v + b1); out = (1 + tanh(w2.h + b2)) / 2.
// ------------------------------------------------------------------------------------------
constexpr int kValueGroup = 4;

__global__ void __launch_bounds__(256) value_head_kernel(const __half* __restrict__ act, int chunk_rows, int c_in,
                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                         const float* __restrict__ ip1_wt /*[361][hidden]*/,
                                                         const float* __restrict__ ip1_b, int hidden,
                                                         const float* __restrict__ ip2_w, const float* __restrict__ ip2_b,
                                                         int n, float* __restrict__ winrate) {
    extern __shared__ float hs[];
    float* w_s = hs;                              // [9][c_in]
    float* v_s = w_s + 9 * c_in;                  // [G][361]
    float* h_s = v_s + kValueGroup * kPoints;     // [G][hidden]
    const int tid = threadIdx.x;
    const int pos0 = blockIdx.x * kValueGroup;
    for (int i = tid; i < 9 * c_in; i += blockDim.x) {
        const int t = i / c_in, c = i - t * c_in;
        w_s[i] = w[c * 9 + t];
    }
    __syncthreads();
    for (int i = tid; i < kValueGroup * kPoints; i += blockDim.x) {
        const int g = i / kPoints, p = i - g * kPoints;
        float v = 0.0f;
        if (pos0 + g < n) {
            const int y = p / kBoard, x = p - y * kBoard;
            v = elu1(bias[0] + head_conv_pixel(act, chunk_rows, c_in, w_s, (pos0 + g) * 400, y, x));
        }
        v_s[i] = v;
    }
    __syncthreads();
    for (int o = tid; o < hidden; o += blockDim.x) {
        float a[kValueGroup];
#pragma unroll
        for (int g = 0; g < kValueGroup; g++) a[g] = 0.0f;
        for (int i = 0; i < kPoints; i++) {
            const float wv = ip1_wt[(size_t)i * hidden + o];
#pragma unroll
            for (int g = 0; g < kValueGroup; g++) a[g] = fmaf(wv, v_s[g * kPoints + i], a[g]);
        }
#pragma unroll
        for (int g = 0; g < kValueGroup; g++) h_s[g * hidden + o] = elu1(a[g] + ip1_b[o]);
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    if (warp < kValueGroup && pos0 + warp < n) {
        float a = 0.0f;
        for (int o = lane; o < hidden; o += 32) a = fmaf(ip2_w[o], h_s[warp * hidden + o], a);
        a = warp_sum(a);
        if (lane == 0) winrate[pos0 + warp] = (1.0f + tanhf(a + ip2_b[0])) * 0.5f;
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
cudaError_t launch_expand(const uint32_t* planes, const uint8_t* rotation, int n, __half* x0, int chunk_rows,
                          cudaStream_t st) {
    const int rows = n * 441;
    expand_planes_kernel<<<(rows + 255) / 256, 256, 0, st>>>(planes, rotation, n, x0, chunk_rows);
    return cudaGetLastError();
}

cudaError_t trunk_kernel_setup() {
    return cudaFuncSetAttribute(trunk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTrunkSmemBytes);
}

cudaError_t launch_trunk(const TrunkParams& p, int grid, bool cooperative, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kTrunkThreads);
    cfg.dynamicSmemBytes = kTrunkSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = cooperative ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, trunk_kernel, p);
}

cudaError_t launch_policy_head(const __half* act, int chunk_rows, int c_in, const float* w, const float* bias,
                               const uint8_t* rotation, int n, float temp, float* probs, cudaStream_t st) {
    const size_t smem = (9 * c_in + kPoints) * sizeof(float);
    policy_head_kernel<<<n, 384, smem, st>>>(act, chunk_rows, c_in, w, bias, rotation, temp, probs);
    return cudaGetLastError();
}

cudaError_t launch_value_head(const __half* act, int chunk_rows, int c_in, const float* w, const float* bias,
                              const float* ip1_wt, const float* ip1_b, int hidden, const float* ip2_w,
                              const float* ip2_b, int n, float* winrate, cudaStream_t st) {
    const size_t smem = (9 * c_in + kValueGroup * kPoints + kValueGroup * hidden) * sizeof(float);
    value_head_kernel<<<(n + kValueGroup - 1) / kValueGroup, 256, smem, st>>>(act, chunk_rows, c_in, w, bias, ip1_wt,
                                                                             ip1_b, hidden, ip2_w, ip2_b, n, winrate);
    return cudaGetLastError();
}

}  // namespace lb2
