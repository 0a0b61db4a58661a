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
#include <cuda_fp8.h>
#include <algorithm>
#include <cstdio>

#include "lb2_kernels.cuh"
#include "lb2_ptx.cuh"

namespace lb2 {

// tuning knobs (tools/ab_variants.py builds variants side by side)
#ifndef LB2_PROBE_TAP
#define LB2_PROBE_TAP 4   // tap after which the MMA issuer probes the next stage; < 0: before the stage's MMAs
#endif
#ifndef LB2_RING_SLEEP
#define LB2_RING_SLEEP 64 // ns between polls of the work-item ring by the scout / publisher / stage forwarder
#endif
#ifndef LB2_EPI_GROUP
#define LB2_EPI_GROUP 2   // 8-column TMEM loads the epilogue issues back to back
#endif
constexpr int kEpiGroup = LB2_EPI_GROUP;
#ifndef LB2_PDL
#define LB2_PDL 0         // trunk and heads start under programmatic dependent launch (prologue overlaps the predecessor's tail)
#endif
#ifndef LB2_L2_HINTS
#define LB2_L2_HINTS 3    // L2 eviction priority. 1: activation stores evict_last; 2: + activation loads evict_last; 3: stores evict_last, loads evict_first (measured best: DRAM write-back 285 -> 173 MB per launch)
#endif
#ifndef LB2_RELAXED_HANDBACK
#define LB2_RELAXED_HANDBACK 1   // peer CTA hands the accumulator back with a relaxed remote arrive (no fence behind its stores)
#endif
#ifndef LB2_RELAXED_FORWARD
#define LB2_RELAXED_FORWARD 1    // the peer's "my stage landed" arrive on the leader's barrier carries no fence either: the stage was
                                 // written by TMA and is complete when the peer's own barrier flips — the forwarding thread has
                                 // no writes of its own to publish, and the MMA reads the peer's shared memory itself (no cache
                                 // in between). `.release.cluster` cost a MEMBAR.ALL.GPU round trip per stage: trunk 331 -> 319 us.
#endif
#ifndef LB2_PEER_DIRECT
#define LB2_PEER_DIRECT 1        // resident-weights mode, stages that carry activations only: the peer CTA's TMA loads complete on the
                                 // LEADER's stage barrier (cp.async.bulk.tensor.cta_group::2 with a remote mbarrier) and the peer's
                                 // producer adds its own arrival there — no local barrier, no forwarding thread in between
                                 // (the forward hop was ~0.4 us of the ~1.7 us from issuing a load to the MMA warp seeing it)
#endif
#ifndef LB2_RELOAD_FORWARD_RELEASE
#define LB2_RELOAD_FORWARD_RELEASE 0   // 1: with LB2_PEER_DIRECT, the few stages that are still forwarded (a new job's weights: bulk copies cannot
                                       // signal a remote mbarrier) use the formally ordered .release.cluster arrive. Measured: its fence sits on the
                                       // critical path of every layer switch, trunk 377.5 -> 385.5 us (4 interleaved rounds). Off.
#endif
#ifndef LB2_EPI_PAIR_HALVES
#define LB2_EPI_PAIR_HALVES 1    // ordinary epilogue: unit u = (half u & 1, column block u >> 1): the two units of a group share
                                 // their eight bias values — one shared-memory read (the port the tensor core's operands come
                                 // through) instead of two
#endif
#ifndef LB2_CONSUMER_PROXY_FENCE
#define LB2_CONSUMER_PROXY_FENCE 2   // consumer-side fence.proxy.async (MEMBAR.ALL.GPU, ~0.8 us) between seeing an item's dependency flags and its
                                     // TMA loads: 1 = in the producer, right before the loads; 2 = in the scout warp, right after the flag
                                     // acquires and before it releases the item to the producer (off the load path: the scout runs items
                                     // ahead); 0 = none (the writer's fence before its flag release already orders the proxies)
#endif
// LB2_LITE_SEPARATE_ACC (default 0, lb2_kernels.cuh): lite mode: 1 = the e4m3 correction terms accumulate in TMEM columns of their own (kCorrCols to the right,
                                  // c_out <= 64 only) and the epilogue adds the two sums; 0 = straight onto the fp16 sum. kind::f8f6f4
                                  // adds an addend 2^-24 of the accumulator exactly (tools/mma_mixed_test.cu) and the results are
                                  // the same either way (value max 9.3e-5), so: 0
#ifndef LB2_DISCARD_DEAD
#define LB2_DISCARD_DEAD 0   // 1: the tile publisher drops activation tiles nobody will read again from L2 (discard.global.L2) instead of
                             // letting them be written back. Correct (tests, soak) and useless: DRAM write-back 389 -> 361 MB per launch,
                             // launch time unchanged — with ~105 MB of live activations the lines are evicted (and written back)
                             // long before their last reader is done. Off.
#endif
#ifndef LB2_PACKED_F32
#define LB2_PACKED_F32 1    // epilogue arithmetic in packed fp32 pairs (FFMA2 / FADD2 / FMUL2): the same bits at fewer issue slots
#endif
#ifndef LB2_EPI_PIPE
#define LB2_EPI_PIPE 1    // epilogue keeps the TMEM loads of the next two units in flight
#endif

// timing experiments (results are wrong under them) exist only in builds with -DLB2_DEBUG_KNOBS
#ifdef LB2_DEBUG_KNOBS
#define kDebugFlags(P) ((P).debug_flags)
#else
#define kDebugFlags(P) 0
#endif

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
// expand: one thread per row of the S=21 row space of either net; writes 32 channels = 4 chunks of
// 8 fp16. Spare threads prefetch a read-late buffer into L2.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) expand_planes_kernel(const ExpandArgs A) {
    if (LB2_PDL) grid_dep_launch();   // the trunk may set itself up while we run; it waits for us before reading
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0 && A.sched) {
        // start of an evaluation: new flag epoch, work-item claim counters back to zero. Everything that used the old values
        // precedes this kernel in its stream; the trunk launch(es) of this evaluation follow it. Keeping this state on the
        // device makes a launch sequence independent of its history, i.e. replayable from a CUDA graph.
        A.sched[kSchedEpoch] += 1u;
        A.sched[0] = A.sched[1] = A.sched[2] = 0u;
    }
    const size_t per_net = (size_t)A.n * 441;
    const size_t total = per_net * A.n_nets;
    if (t >= total) {
        const size_t i = (t - total) * 128;
        if (i < A.pf_bytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.pf + i));
        return;
    }
    const int k = (int)(t / per_net);
    const int row = (int)(t - (size_t)k * per_net);
    const int pos = row / 441, rem = row - pos * 441;
    const int y = rem / 21, x = rem - y * 21;
    uint32_t bits = 0;
    if (x < kBoard && y < kBoard) {
        const int sym = A.ensemble ? (pos & 7) : (A.rotation[pos] & 7);
        const int src = rotate_idx(y * kBoard + x, sym);
        bits = A.planes[k][(size_t)(A.ensemble ? pos >> 3 : pos) * kPoints + src];
    }
    const uint32_t one = 0x3C00u;  // fp16 1.0
    __half* x0 = A.x0[k];
    const int chunk_rows = A.chunk_rows[k];
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
//   Warps: 0 = TMA producer, 1 = MMA issuer (+ TMEM owner), 2..9 = epilogue (warp w reads TMEM
//   lane quadrant w%4 and column half (w-2)/4).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int n_tap_groups(int ksize) { return ksize == 3 ? 1 : 3; }
__device__ __forceinline__ int tap_group_begin(int g) { return g == 0 ? 0 : (g == 1 ? 9 : 17); }
__device__ __forceinline__ int tap_group_end(int ksize, int g) { return ksize == 3 ? 9 : (g == 0 ? 9 : (g == 1 ? 17 : 25)); }

// Producer items [lo, hi] of J.dep_job whose outputs overlap the input rows (tile + halo) of `tile`.
__device__ __forceinline__ void dependency_range(const LayerJob& J, int tile, int& lo, int& hi) {
    const int r_lo = tile * kTileRows - J.halo;
    const int r_hi = tile * kTileRows + kTileRows - 1 + J.halo;
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
    if (J.group_tiles) {
        // Position groups: the rows behind a group's last position belong to the next group, which is scheduled later. Only
        // outputs in this group's padding rows would read them, so the tile need not (and, in-order claiming, must not) wait.
        const int cap = (tile / J.group_tiles + 1) * J.dep_group_tiles - 1;
        if (hi > cap) hi = cap;
    }
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// timeline tracing (debug): event e of this CTA's `it`-th item
#define LB2_TRACE(it, e)                                                                             \
    do {                                                                                             \
        if (P.trace && (it) < (uint32_t)kTraceItems)                                                 \
            P.trace[((size_t)blockIdx.x * kTraceItems + (it)) * kTraceEvents + (e)] = global_ns();   \
    } while (0)

// two floats -> two e4m3 (round to nearest even, saturating at +-448) in the low 16 bits: lo = first, hi = second
__device__ __forceinline__ uint32_t cvt_e4m3x2(float first, float second) {
    uint16_t r;
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(r) : "f"(second), "f"(first));
    return (uint32_t)r;
}

// ELU(alpha = 1): v > 0 ? v : exp(v) - 1, written branch-free as max(v, min(exp(v) - 1, 0)).
__device__ __forceinline__ float elu_fast(float v) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * 1.4426950408889634f));  // inf for large v is clamped by the min
    return fmaxf(v, fminf(e - 1.0f, 0.0f));  // v>0: e-1>0 -> max(v,0)=v; v<=0: e-1 in (-1,0] and e-1 >= v
}

// the same for two values with packed fp32 instructions (FMUL2, FADD2): identical bits, fewer issue slots
__device__ __forceinline__ void elu_fast2(float& a, float& b) {
    const f32x2 t = mul2(pack2(a, b), pack2(1.4426950408889634f, 1.4426950408889634f));
    float ta, tb, ea, eb;
    unpack2(t, ta, tb);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ea) : "f"(ta));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(eb) : "f"(tb));
    float ma, mb;
    unpack2(add2(pack2(ea, eb), pack2(-1.0f, -1.0f)), ma, mb);
    a = fmaxf(a, fminf(ma, 0.0f));
    b = fmaxf(b, fminf(mb, 0.0f));
}

// Item index -> (job, item within the job). Rounds are consecutive in the launch-wide order; inside
// a round the items of its jobs (policy layer l — possibly as two column-split jobs — and value
// layer l) are dealt round robin, so that at any time every cluster works on a mix of wide/long and
// narrow/short items instead of the whole GPU switching between regimes. Cycle c of the deal gives
// one item to every job that still has more than c. `r` is the caller's monotonically advancing
// round cursor. Only the scout warp runs this, once per item.
__device__ __forceinline__ void locate_item(const TrunkParams& P, const LayerJob* jobs, int q, int& r, int& job, int& idx) {
    while (q >= P.round_base[r + 1]) r++;
    int local = q - P.round_base[r];
    const int first = P.round_first[r], m = P.round_jobs[r];
    if (m == 1) { job = first; idx = P.round_idx0[r] + local; return; }   // net-major order: a slice of one job
    int done = 0;   // cycles dealt so far
    for (;;) {
        int active = 0, shortest = 0x7fffffff;
        for (int k = 0; k < m; k++) {
            const int n = jobs[first + k].n_items;
            if (n > done) { active++; shortest = n < shortest ? n : shortest; }
        }
        const int phase = active * (shortest - done);   // items dealt until the shortest active job runs out
        if (local < phase) {
            idx = done + local / active;
            int nth = local % active;
            for (int k = 0; k < m; k++)
                if (jobs[first + k].n_items > done && nth-- == 0) { job = first + k; return; }
        }
        local -= phase;
        done = shortest;
    }
}

constexpr int kClaimRing = 16;  // work-item ring entries per CTA (claimed-but-unpublished items)
#ifndef LB2_CLAIM_AHEAD
#define LB2_CLAIM_AHEAD 4
#endif
constexpr int kClaimAhead = LB2_CLAIM_AHEAD;  // how far ahead of the last published item the scout may hand out items
constexpr uint32_t kEndJob = 63;  // job field of the ring entry that ends a CTA's walk

// Work-item ring. The scout warp of the cluster leader decides which item the cluster processes
// k-th (static: round robin over the launch's item list; dynamic: the next one of a global in-order
// counter — greedy list scheduling, still strictly increasing per cluster), resolves it to
// (job, index within the job) ONCE and writes a 32-bit entry [tag:5 | job:6 | index:21] into the
// ring of every CTA of the cluster. All other roles just read entry k: one shared-memory load,
// no search through the round table, no acquire fence (tag and payload share the word).
__device__ __forceinline__ uint32_t item_tag(uint32_t k) { return ((k / kClaimRing) & 15u) + 1u; }   // 1..16, never 0
static_assert(kMaxLaunchJobs < (int)kEndJob, "job field of the item ring");
__device__ __forceinline__ uint32_t item_pack(uint32_t k, uint32_t job, uint32_t idx) {
    return (item_tag(k) << 27) | (job << 21) | idx;
}
__device__ __forceinline__ bool item_peek(const uint32_t* ring, uint32_t k, uint32_t& v) {
    v = ld_volatile_shared(ring + k % kClaimRing);
    return (v >> 27) == item_tag(k);
}
// blocking read of entry k; false at the end marker. `relaxed`: roles off the critical path back off
// between polls so their spinning stays off the shared-memory port the tensor core reads through.
template <bool relaxed = false>
__device__ __forceinline__ bool item_get(const uint32_t* ring, uint32_t k, int& job, int& idx) {
    uint32_t v;
    while (!item_peek(ring, k, v)) { if (relaxed && LB2_RING_SLEEP) __nanosleep(LB2_RING_SLEEP); }
    job = (int)((v >> 21) & 63u);
    idx = (int)(v & 0x1fffffu);
    return job != (int)kEndJob;
}

// The MMA issuer's own ring: everything it needs to know about item k in one word
// [tag:5 | end:1 | 5x5:1 | halo/8:4 | virtual slabs:6 | c_out/8:6 | S == 21:1 | kind::f16 slabs:7], so that a single
// shared-memory load separates the last MMA of one item from the first of the next.
__device__ __forceinline__ uint32_t geom_pack(uint32_t k, const LayerJob* J) {
    if (!J) return (item_tag(k) << 27) | (1u << 25);
    return (item_tag(k) << 27) | ((J->ksize == 5 ? 1u : 0u) << 24) | ((uint32_t)(J->halo >> 3) << 20) |
           ((uint32_t)J->n_slabs << 14) | ((uint32_t)(J->n_out >> 3) << 8) | ((J->S == 21 ? 1u : 0u) << 7) | (uint32_t)J->n_f16_slabs;
}

// kRes (pairs only): resident-weights mode, see lb2_kernels.cuh. Shared memory:
//   streaming:  [stages: A slab + B block][ctrl][bias][head weights][job table]
//   resident:   [this CTA's half of the layer's weights][stages: A slab only][ctrl][bias][head weights][job table]
// kOutModes: which LayerJob::out_mode values besides kOutPlain may occur in the launch (bit kOutLo16 - 1: precise, bit
// kOutFp8 - 1: lite) — the epilogue's extra stores are compiled in only where a launch can need them.
// Slots of the shared-memory stage ring a job uses. Streaming modes: all of them, at a fixed slot size. Resident-weights
// mode: the slots hold activation slabs only and are packed at the job's own slab size — 6 slots of 304 rows for the 3x3
// layers (halo 24), 5 of 352 rows for the 5x5 first layer (halo 48) in the same 58 KB. The sixth slot matters: a slot is
// free only when its MMAs have COMPLETED, about two stages behind the issuer, so of five slots only three were ahead of it
// and the first slab of every item arrived ~0.4 us late (1.7 us from issuing a load to the issuer seeing it).
template <bool kRes, int kStages>
__device__ __forceinline__ int ring_slots(int halo) { return kRes ? (halo <= kResHalo3 ? kStagesRes3 : kStagesRes) : kStages; }

template <bool kPair, bool kRes, int kOutModes>
__global__ void __launch_bounds__(kTrunkThreads, 1) trunk_kernel(const __grid_constant__ TrunkParams P) {
    static_assert(!kRes || kPair, "resident weights need CTA pairs");
    constexpr bool kLo16 = (kOutModes & 1) != 0, kFp8 = (kOutModes & 2) != 0;
    constexpr int kStages = kRes ? kStagesRes : (kPair ? kStagesPair : kStagesSingle);   // (kRes: see ring_slots)
    constexpr int kStageBytes = kRes ? kASlabBytes : (kPair ? kStageBytesPair : kStageBytesSingle);
    constexpr int kRingOff = kRes ? kResWeightBytes : 0;
    constexpr int kCtrlOff = kRes ? kResWeightBytes + kResRingBytes : kTrunkRingBytes;
    constexpr int kJobSlots = kRes ? kResJobs : kMaxLaunchJobs;
    constexpr int kHeadW = kRes ? kResHeadSlots : kHeadSlots;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* ring = smem + kRingOff;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kCtrlOff);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* tfull_bar = empty_bar + kMaxStages;  // [2] accumulator ready   (MMA -> epilogue)
    uint64_t* tempty_bar = tfull_bar + 2;          // [2] accumulator drained (epilogue -> MMA)
    uint64_t* pub_bar = tempty_bar + 2;            // [kPubDepth] tile stored (epilogue -> publisher)
    uint64_t* pfull_bar = pub_bar + kPubDepth;     // [kStages] pair mode: the peer CTA's stage landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pfull_bar + kMaxStages);
    volatile uint32_t* pub_done = tmem_slot + 1;   // tiles published so far by this CTA
    uint32_t* deps_ready = tmem_slot + 2;          // items whose dependencies the scout warp has seen satisfied
    uint32_t* item_ring = tmem_slot + 4;           // [kClaimRing] work-item entries (see item_pack)
    uint32_t* geom_ring = item_ring + kClaimRing;  // [kClaimRing] the MMA issuer's view of them (see geom_pack)
    // broadcast reads (one wavefront per warp-wide LDS.128), resident for the whole launch
    float* bias_all = reinterpret_cast<float*>(smem + kCtrlOff + kCtrlBytes);   // [kJobSlots][128]
    float* headw_all = bias_all + kJobSlots * 128;                              // [kHeadW][9][128]
    LayerJob* jobs_s = reinterpret_cast<LayerJob*>(headw_all + kHeadW * 9 * 128);    // [kJobSlots] job table copy

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // Pair mode (cta_group::2): an item is two adjacent 256-row tiles; the CTA with cluster rank r
    // owns tile 2*item + r. Single mode: item == tile.
    const int rank = kPair ? (int)cluster_ctarank() : 0;
    const bool leader = (rank == 0);
    auto tile_of = [&](int idx) { return kPair ? 2 * idx + rank : idx; };

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 127u) __trap();  // TMA destinations need 128-byte alignment
        // pair mode: the leader's "stage full" barrier also counts the peer's "my half landed" arrive
        for (int i = 0; i < kMaxStages; i++) { mbar_init(full_bar + i, (kPair && leader) ? 2 : 1); mbar_init(empty_bar + i, 1); }
        for (int i = 0; i < 2; i++) { mbar_init(tfull_bar + i, 1); mbar_init(tempty_bar + i, kPair ? 2 * kEpilogueWarps : kEpilogueWarps); }
        for (int i = 0; i < kPubDepth; i++) mbar_init(pub_bar + i, kEpilogueWarps);
        *pub_done = 0;
        *deps_ready = 0;
        for (int i = 0; i < 2 * kClaimRing; i++) item_ring[i] = 0;
        fence_mbar_init();
        fence_proxy_async_smem();
        for (int i = 0; i < kMaxTensorMaps; i++) tma_prefetch_desc(&P.tmaps[i]);
    }
    if (warp == 1) { if (kPair) tmem_alloc_pair<512>(tmem_slot); else tmem_alloc<512>(tmem_slot); }
    // the job table, every job's bias and both fused-head weight sets stay resident in smem for the
    // whole launch (the per-item descriptor reads would otherwise cost an L2 round trip per role)
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(P.jobs);
        uint32_t* dst = reinterpret_cast<uint32_t*>(jobs_s);
        for (int i = threadIdx.x; i < P.n_jobs * (int)(sizeof(LayerJob) / 4); i += blockDim.x) dst[i] = src[i];
    }
    const LayerJob* jobs = P.jobs;
    for (int i = threadIdx.x; i < P.n_jobs * 128; i += blockDim.x) {
        const int jj = i >> 7, c = i & 127;
        bias_all[i] = c < jobs[jj].n_out ? jobs[jj].bias[c] : 0.0f;
    }
    for (int jj = 0; jj < P.n_jobs; jj++) {
        if (!jobs[jj].head_taps) continue;
        float* dst = headw_all + jobs[jj].head_slot * (9 * 128);
        for (int i = threadIdx.x; i < 9 * jobs[jj].n_out; i += blockDim.x) dst[(i / jobs[jj].n_out) * 128 + i % jobs[jj].n_out] = jobs[jj].head_w[i];
    }
    tc_fence_before_sync();
    __syncthreads();
    if (kPair) cluster_sync_all();  // the peer's barriers and ring are initialised before anyone signals them
    tc_fence_after_sync();
    if (LB2_PDL) {
        // everything above touched only launch-invariant data (weights, job table); the activations,
        // flags and the item counter belong to the predecessor in the stream until it has completed
        grid_dep_wait();
        grid_dep_launch();
    }
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t epoch = P.sched[kSchedEpoch];   // set by the expand kernel in front of this launch
    jobs = jobs_s;  // from here on every role reads the shared-memory copy
    if (P.trace && threadIdx.x == 0) {  // effective SM clock: cycle and wall timestamps at both ends
        unsigned long long* t = P.trace + ((size_t)blockIdx.x * kTraceItems + (kTraceItems - 1)) * kTraceEvents;
        t[0] = global_ns(); t[1] = (unsigned long long)clock64();
    }

    if (warp == 0) {
        // ================================ TMA producer ================================
        // One lane walks the item ring and issues the copies. In pair mode each CTA loads the A slab of its own tile and its
        // half of the output channels of the B block. Ring slots carry their own phase parity (bit k of `pbits` = parity of
        // slot k's next fill), so the number of slots in use may change from job to job (ring_slots).
        if (lane == 0) {
        int st = 0; uint32_t pbits = 0; uint32_t pit = 0;
        int ring_n = 0;          // slots of the ring geometry in use (0: none yet)
        int resident_job = -1;   // kRes: the job whose weights (this CTA's half) are in shared memory
        const uint64_t ld_policy = LB2_L2_HINTS == 2 ? l2_policy_evict_last() : (LB2_L2_HINTS == 3 ? l2_policy_evict_first() : 0);
        const uint32_t leader_full0 = kPair ? mapa_u32(full_bar, 0) : 0u;   // the leader's full_bar[0] as a shared::cluster address
        for (int jj, idx; item_get(item_ring, pit, jj, idx); pit++) {
            const LayerJob& J = jobs[jj];
            const int tile = tile_of(idx);
            const int halo = J.halo, ksize = J.ksize, n_out = J.n_out, n_slabs = J.n_slabs, tmap = J.tmap, n_real = J.n_real_slabs;
            const int tb0 = J.term_base[0], tb1 = J.term_base[1], tb2 = J.term_base[2];
            LB2_TRACE(pit, 0);
            if (P.trace && pit < (uint32_t)kTraceItems) P.trace[((size_t)blockIdx.x * kTraceItems + pit) * kTraceEvents + 15] = (unsigned long long)((jj << 21) | idx);
            // the scout warp polls the dependency flags ahead of us; wait for its go-ahead
            while (ld_acquire_cta_shared(deps_ready) <= pit) {}
            const int rows_halo = kTileRows + 2 * halo;
            const uint32_t a_bytes = rows_halo * 32;
            const int row0_8 = (tile * kTileRows - halo) / 8;  // exact: both multiples of 8
            const int ng = n_tap_groups(ksize);
            const int n_mine = kPair ? (n_out >> 1) : n_out;   // output channels whose weights this CTA stages
            const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(kPair ? J.wpk2 : J.wpk);
            const int rn = ring_slots<kRes, kStages>(halo);
            const uint32_t sbytes = kRes ? a_bytes : (uint32_t)kStageBytes;   // kRes: slots are packed at the job's own slab size
            // kRes: a new job's weights replace the resident ones unit by unit (unit = the taps of one
            // stage), each riding on its stage's barrier. Overwriting unit u is safe once the MMAs that
            // read the bytes under it are done: for a job of the same geometry with at least as many stages as the ring has
            // slots that is implied by owning the stage's ring slot (the old unit u was last read a whole ring ago); otherwise
            // — and whenever the ring's slot size changes — drain the pipeline first.
            const bool reload = kRes && jj != resident_job;
            bool drain = ring_n != 0 && rn != ring_n;
            if (reload && resident_job >= 0) {
                const LayerJob& O = jobs[resident_job];
                drain |= !(O.ksize == ksize && O.n_out == n_out && O.n_slabs == n_slabs && n_slabs * ng >= rn);
            }
            LB2_TRACE(pit, 1);
            if (LB2_CONSUMER_PROXY_FENCE == 1) fence_proxy_async();  // order the TMA (async proxy) reads after the acquires above
            if (drain)
                for (int k = 0; k < ring_n; k++) mbar_wait(empty_bar + k, ((pbits >> k) & 1u) ^ 1u);   // the wait the slot's next fill would do, done now
            if (rn != ring_n) { st = 0; ring_n = rn; }
            uint32_t unit_off = 0;   // kRes: byte offset of the stage's weights inside the resident area
            int term = 0, ts = 0;    // virtual slab s = term * n_real + ts
            // peer CTA, activations only: straight onto the leader's barrier (see LB2_PEER_DIRECT)
            const bool direct = LB2_PEER_DIRECT && kRes && !leader && !reload;
            for (int s = 0; s < n_slabs; s++) {
                for (int g = 0; g < ng; g++) {
                    const uint32_t b_bytes = (tap_group_end(ksize, g) - tap_group_begin(g)) * n_mine * 32;
                    mbar_wait(empty_bar + st, ((pbits >> st) & 1u) ^ 1u);
                    pbits ^= 1u << st;
                    uint8_t* sa = ring + st * sbytes;
                    const bool skip_b = (kDebugFlags(P) & 8) != 0 || (kRes && !reload), skip_a = (kDebugFlags(P) & 16) != 0;
                    // split-operand modes: every term of the K loop has its own set of input chunk planes
                    const int ac = (term == 0 ? tb0 : (term == 1 ? tb1 : tb2)) + 2 * ts;
                    if (direct) {
                        const uint32_t lbar = leader_full0 + 8u * (uint32_t)st;
                        mbar_arrive_expect_tx_cluster_relaxed(lbar, a_bytes);
                        if (LB2_L2_HINTS >= 2) tma_load_3d_pair_hint(sa, &P.tmaps[tmap], lbar, 0, row0_8, ac, ld_policy);
                        else tma_load_3d_pair(sa, &P.tmaps[tmap], lbar, 0, row0_8, ac);
                    } else {
                        mbar_arrive_expect_tx(full_bar + st, (skip_a ? 0u : a_bytes) + (skip_b ? 0u : b_bytes));
                        if (!skip_a) {
                            if (LB2_L2_HINTS >= 2) tma_load_3d_hint(sa, &P.tmaps[tmap], full_bar + st, 0, row0_8, ac, ld_policy);
                            else tma_load_3d(sa, &P.tmaps[tmap], full_bar + st, 0, row0_8, ac);
                        }
                        if (!skip_b) bulk_load_1d(kRes ? smem + unit_off : sa + kASlabBytes, wsrc + (kPair ? rank * b_bytes : 0u), b_bytes, full_bar + st);
                        wsrc += kPair ? 2 * b_bytes : b_bytes;
                        unit_off += b_bytes;
                    }
                    if (++st == rn) st = 0;
                    if (s == 0 && g == 0) LB2_TRACE(pit, 2);
                }
                if (++ts == n_real) { ts = 0; term++; }
            }
            LB2_TRACE(pit, 3);
            resident_job = jj;
        }
        }
    } else if (warp == 1 && !leader) {
        // ================================ pair mode, peer CTA ==========================
        // No MMAs are issued here (the leader's tcgen05.mma.cta_group::2 drives both SMs); this
        // warp only tells the leader when each of OUR stages has landed.
        if (lane == 0) {
            // (with LB2_PEER_DIRECT only the stages of an item that brings a new job's weights use our own barriers: the parity
            // of each is kept per slot)
            int stage = 0; uint32_t parity = 0; uint32_t fit = 0;
            int resident_job = -1, ring_n = 0;
            for (int jj, idx; item_get<true>(item_ring, fit, jj, idx); fit++) {
                const int n_st = jobs[jj].n_slabs * n_tap_groups(jobs[jj].ksize);
                const int rn = ring_slots<kRes, kStages>(jobs[jj].halo);
                if (rn != ring_n) { stage = 0; ring_n = rn; }
                const bool direct = LB2_PEER_DIRECT && kRes && jj == resident_job;
                resident_job = jj;
                if (direct) { stage = (stage + n_st) % rn; continue; }
                for (int s = 0; s < n_st; s++) {
                    mbar_wait(full_bar + stage, (parity >> stage) & 1u);
                    parity ^= 1u << stage;
                    // (resident mode with direct signalling: only the few stages that bring a new job's weights come through here —
                    // they take the formally ordered .release.cluster arrive; the streaming modes forward every stage and keep the
                    // relaxed one, see LB2_RELAXED_FORWARD)
                    if (LB2_RELAXED_FORWARD && !(LB2_RELOAD_FORWARD_RELEASE && LB2_PEER_DIRECT && kRes)) mbar_arrive_remote_relaxed(full_bar + stage, 0); else mbar_arrive_remote(full_bar + stage, 0);
                    if (++stage == rn) stage = 0;
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        // One elected lane runs the whole pipeline. The tensor pipe only queues a few MMAs, so the
        // time between the last MMA of one stage (or item) and the first of the next must stay
        // short: the geometry of the NEXT item is fetched while this item's MMAs are in flight,
        // and the readiness of the NEXT stage is probed (non-blocking) in the middle of this
        // stage's MMAs, so both latencies hide behind queued tensor work.
        if (elect_one()) {
            int stage = 0; uint32_t cbits = 0; uint32_t it = 0;   // bit k of cbits: parity of ring slot k's next use
            int ring_n = 0;
            bool next_ready = false;  // full_bar[stage] already observed complete for its next use
            for (;; it++) {
                uint32_t gw;
                while (!item_peek(geom_ring, it, gw)) {}
                if (gw & (1u << 25)) break;
                const int ksize = (gw & (1u << 24)) ? 5 : 3, halo = (int)((gw >> 20) & 15u) << 3;
                const int n_slabs = (int)((gw >> 14) & 63u), n_out = (int)((gw >> 8) & 63u) << 3;
                const int S = (kDebugFlags(P) & 2) ? 0 : ((gw & 128u) ? 21 : 20), DX = (kDebugFlags(P) & 2) ? 0 : 1;
                const int n_f16 = (int)(gw & 127u);   // 3x3 layers: slabs from this one on are e4m3 (K = 32) MMAs
                const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
                LB2_TRACE(it, 4);
                mbar_wait(tempty_bar + acc, acc_phase ^ 1);
                LB2_TRACE(it, 5);
                const uint32_t idesc = umma_idesc_f16(kPair ? 256 : 128, n_out);
                const int rows_halo = kTileRows + 2 * halo;
                const int n_mine = kPair ? (n_out >> 1) : n_out;   // B rows held per CTA
                // descriptor = hi32 (SBO = 128 B, version 1) : lo32 (LBO << 16 | start address >> 4)
                const uint64_t desc_hi = static_cast<uint64_t>((128u >> 4) | (1u << 14)) << 32;
                const uint32_t a_lo_base = ((uint32_t)(rows_halo * 16) >> 4) << 16;
                const uint32_t b_lo_base = ((uint32_t)(n_mine * 16) >> 4) << 16;
                const uint32_t b_step = (uint32_t)(n_mine * 32) >> 4;  // one tap of B, in 16-byte units
                const uint32_t d0 = tmem_base + (acc * 2 + 0) * 128;
                const uint32_t d1 = tmem_base + (acc * 2 + 1) * 128;
                const int ng = n_tap_groups(ksize);
                const int pad = ksize >> 1;
                const int n_st = n_slabs * ng;
                const int rn = ring_slots<kRes, kStages>(halo);
                const uint32_t sbytes = kRes ? (uint32_t)rows_halo * 32u : (uint32_t)kStageBytes;
                if (rn != ring_n) { stage = 0; ring_n = rn; next_ready = false; }   // (the producer drained the ring before it changed the slot size)
                uint32_t accumulate = 0;
                int g = 0;
                uint32_t res16 = 0;   // kRes: offset of the stage's weights in the resident area, in 16-byte units
                for (int s = 0; s < n_st; s++) {
                    if (!next_ready) mbar_wait(full_bar + stage, (cbits >> stage) & 1u);
                    cbits ^= 1u << stage;
                    tc_fence_after_sync();
                    if (s == 0) LB2_TRACE(it, 6);
                    const int nstage = stage + 1 == rn ? 0 : stage + 1;
                    const uint32_t nphase = (cbits >> nstage) & 1u;
                    if (LB2_PROBE_TAP < 0) next_ready = mbar_try_wait(full_bar + nstage, nphase);
                    const uint32_t a_addr = smem_u32(ring + stage * sbytes);
                    // (address of row `halo` of the A slab) >> 4; a tap shifts it by dy*S+dx rows
                    const uint32_t a16 = (a_addr >> 4) + halo;
                    // B: behind the A slab of the stage, or (kRes) this stage's unit of the resident weights
                    uint32_t b16 = kRes ? (smem_u32(smem) >> 4) + res16 : (a_addr + kASlabBytes) >> 4;
                    if (kRes) res16 += (uint32_t)(ksize == 3 ? 9 : tap_group_end(ksize, g) - tap_group_begin(g)) * b_step;
                    if (ksize == 3 && kFp8 && s >= n_f16) {
                        // lite mode's correction slab: same operand geometry in bytes (16-byte rows = 16 e4m3 channels),
                        // K = 32 = [a8 | lo8] against [Wl8 ; W8], onto the fp16 sum in the same fp32 accumulator (or, see
                        // LB2_LITE_SEPARATE_ACC, into columns of its own)
                        const uint32_t acc_c = (!LB2_LITE_SEPARATE_ACC || s > n_f16) ? 1u : 0u;
                        constexpr uint32_t kCorr = LB2_LITE_SEPARATE_ACC ? kCorrCols : 0;
#pragma unroll
                        for (int t = 0; t < 9; t++) {
                            const int off = (t / 3 - 1) * S + (t % 3 - 1) * DX;
                            const uint64_t bdesc = desc_hi | (b_lo_base | (b16 & 0x3FFFu));
                            const uint32_t a0 = a16 + off;
                            umma_f8<kPair>(d0 + kCorr, desc_hi | (a_lo_base | (a0 & 0x3FFFu)), bdesc, idesc, t ? 1u : acc_c);
                            umma_f8<kPair>(d1 + kCorr, desc_hi | (a_lo_base | ((a0 + 128) & 0x3FFFu)), bdesc, idesc, t ? 1u : acc_c);
                            b16 += b_step;
                            if (t == LB2_PROBE_TAP) next_ready = mbar_test_wait(full_bar + nstage, nphase);
                        }
                    } else if (ksize == 3) {
#pragma unroll
                        for (int t = 0; t < 9; t++) {
                            const int off = (t / 3 - 1) * S + (t % 3 - 1) * DX;
                            const uint64_t bdesc = desc_hi | (b_lo_base | (b16 & 0x3FFFu));
                            const uint32_t a0 = a16 + off;
                            umma_f16<kPair>(d0, desc_hi | (a_lo_base | (a0 & 0x3FFFu)), bdesc, idesc, accumulate);
                            if (!(kDebugFlags(P) & 32))
                                umma_f16<kPair>(d1, desc_hi | (a_lo_base | ((a0 + 128) & 0x3FFFu)), bdesc, idesc, accumulate);
                            accumulate = 1;
                            b16 += b_step;
                            // probe the following stage while the tensor pipe has work queued
                            if (t == LB2_PROBE_TAP) next_ready = mbar_test_wait(full_bar + nstage, nphase);
                        }
                    } else {
                        const int t0 = tap_group_begin(g), t1 = tap_group_end(ksize, g);
                        int kr = t0 / 5, kc = t0 - kr * 5;
                        for (int t = t0; t < t1; t++) {
                            const int off = (kr - pad) * S + (kc - pad) * DX;
                            const uint64_t bdesc = desc_hi | (b_lo_base | (b16 & 0x3FFFu));
                            const uint32_t a0 = a16 + off;
                            umma_f16<kPair>(d0, desc_hi | (a_lo_base | (a0 & 0x3FFFu)), bdesc, idesc, accumulate);
                            umma_f16<kPair>(d1, desc_hi | (a_lo_base | ((a0 + 128) & 0x3FFFu)), bdesc, idesc, accumulate);
                            accumulate = 1;
                            b16 += b_step;
                            if (++kc == 5) { kc = 0; kr++; }
                            if (t == t0 + LB2_PROBE_TAP) next_ready = mbar_test_wait(full_bar + nstage, nphase);
                        }
                        if (++g == ng) g = 0;
                    }
                    umma_commit<kPair>(empty_bar + stage);  // frees the smem stage (in both CTAs) when these MMAs finish
                    stage = nstage;
                }
                umma_commit<kPair>(tfull_bar + acc);  // accumulator complete -> epilogue (both CTAs)
                LB2_TRACE(it, 7);
            }
        }
        __syncwarp();
    } else if (warp < 2 + kEpilogueWarps) {
        // ================================ epilogue ====================================
        // Warp w drains TMEM lane quadrant w%4 (32 rows per 128-row half tile) and one part of the
        // output channels, 8 columns (one 16-byte chunk row) at a time: bias + ELU, then either the
        // fp16 store or the fused head's partial dot products. kEpiGroup TMEM loads are issued back
        // to back; with LB2_EPI_PIPE the next group's loads are in flight while this one is processed.
        const int ew = warp - 2;
        const uint64_t st_policy = LB2_L2_HINTS ? l2_policy_evict_last() : 0;
        const int quad = warp & 3;        // TMEM lane quadrant this warp may read
        const int part = ew >> 2;         // which part of the output channels
        constexpr int G = kEpiGroup;
        uint32_t it = 0;
        for (int jj, idx; item_get(item_ring, it, jj, idx); it++) {
            const LayerJob& J = jobs[jj];
            const int tile = tile_of(idx);
            const int n_out = J.n_out, chunk_rows = J.out_chunk_rows, n_pos = J.n_pos, lo_chunks = kOutModes ? J.lo_chunks : 0;
            const int out_mode = kOutModes ? J.out_mode : kOutPlain;
            const float sc = kOutModes ? J.acc_scale : 1.0f;   // packed weights of the split-operand modes carry a power-of-two scale
            // lite jobs: the e4m3 correction terms sit in an accumulator of their own, kCorrCols columns to the right
            const bool has_corr = LB2_LITE_SEPARATE_ACC && kFp8 && J.n_f16_slabs < J.n_slabs;

            const bool head = J.head_taps != 0, remap = J.remap != 0, wide = (J.S == 21);
            __half* __restrict__ out = J.out;
            float* __restrict__ zbuf = J.zbuf;
            const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
            const float* bs = bias_all + jj * 128;
            const float* headw_s = headw_all + J.head_slot * (9 * 128);
            const int cols = n_out / kColParts;   // columns handled by this warp
            const int col0 = part * cols;
            const int upc = cols >> 3;            // 8-column units per 128-row half tile
            const int n_units = (kDebugFlags(P) & 64) ? 0 : 2 * upc;
            // this thread's row in each of the two half tiles
            int out_row2[2]; bool valid2[2], skip2[2];   // skip: a row of the S = 21 space with no counterpart in the S = 20 space
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int row = tile * kTileRows + h * 128 + quad * 32 + lane;
                int pos, y, x;
                if (wide) { pos = row / 441; const int rem = row - pos * 441; y = rem / 21; x = rem - y * 21; }
                else      { pos = row / 400; const int rem = row - pos * 400; y = rem / 20; x = rem - y * 20; }
                valid2[h] = (x < kBoard) && (y < kBoard) && (!remap || pos < n_pos);
                out_row2[h] = remap ? pos * 400 + y * 20 + x : row;
                // Layer 1 re-addresses S = 21 rows into the S = 20 space: (x, y) <= 19 exist there, and the ones with x == 19 or
                // y == 19 are its zero padding — written here as well (as every other layer does for its own padding rows), so
                // that no launch relies on zeros an earlier launch left behind (dead activation tiles are discarded from L2).
                skip2[h] = remap && (x > kBoard || y > kBoard || pos >= n_pos);
            }
            // plane p, row r of a [planes][chunk_rows] buffer, in rows — 32-bit: planes x rows stays below 2^32 for every batch the
            // host accepts (ensure_workspace), and the 64-bit sign-extending index arithmetic cost the epilogue registers it spilled
            auto row_off = [&](int plane, int r) { return (uint32_t)plane * (uint32_t)chunk_rows + (uint32_t)r; };
            const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 256 + col0;
            // unit -> (half, column block). The fused-head path walks half-major; the ordinary path pairs the halves
            auto unit_addr = [&](int u) { const int h = u >= upc ? 1 : 0; return tbase + h * 128 + (u - h * upc) * 8; };
            auto unit_half = [&](int u) { return LB2_EPI_PAIR_HALVES ? (u & 1) : (u >= upc ? 1 : 0); };
            auto unit_col = [&](int u) { return LB2_EPI_PAIR_HALVES ? (u >> 1) * 8 : (u - (u >= upc ? upc : 0)) * 8; };
            auto unit_addr2 = [&](int u) { return tbase + unit_half(u) * 128 + unit_col(u); };

            // bias + ELU of one unit
            auto activate_b = [&](const uint32_t (&r)[8], uint32_t taddr, const float4& b0, const float4& b1, float (&v)[8]) {
                if (kOutModes) {   // x * 1 + b rounds exactly like x + b: plain jobs of such a launch stay bit-identical
                    float a[8];
#pragma unroll
                    for (int e = 0; e < 8; e++) a[e] = __uint_as_float(r[e]);
                    if (has_corr) {   // (the wait also covers the caller's loads in flight: harmless)
                        uint32_t c[8];
                        tmem_ld_32x8(taddr + kCorrCols, c);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 8; e++) a[e] += __uint_as_float(c[e]);
#ifdef LB2_DEBUG_KNOBS
                        if ((kDebugFlags(P) & 128) && blockIdx.x == 0 && warp == 2 && lane == 5 && it < 40)
                            printf("it %u job %d: main %g corr %g sc %g n_f16 %d n_slabs %d\n", it, jj, __uint_as_float(r[0]), __uint_as_float(c[0]), sc, (int)J.n_f16_slabs, J.n_slabs);
#endif
                    }
#if LB2_PACKED_F32
                    const f32x2 sc2 = pack2(sc, sc);
                    f32x2 p[4] = {fma2(pack2(a[0], a[1]), sc2, pack2(b0.x, b0.y)), fma2(pack2(a[2], a[3]), sc2, pack2(b0.z, b0.w)),
                                  fma2(pack2(a[4], a[5]), sc2, pack2(b1.x, b1.y)), fma2(pack2(a[6], a[7]), sc2, pack2(b1.z, b1.w))};
#pragma unroll
                    for (int e = 0; e < 4; e++) unpack2(p[e], v[2 * e], v[2 * e + 1]);
#else
                    v[0] = fmaf(a[0], sc, b0.x); v[1] = fmaf(a[1], sc, b0.y); v[2] = fmaf(a[2], sc, b0.z); v[3] = fmaf(a[3], sc, b0.w);
                    v[4] = fmaf(a[4], sc, b1.x); v[5] = fmaf(a[5], sc, b1.y); v[6] = fmaf(a[6], sc, b1.z); v[7] = fmaf(a[7], sc, b1.w);
#endif
                } else {
#if LB2_PACKED_F32
                    f32x2 p[4] = {add2(pack2(__uint_as_float(r[0]), __uint_as_float(r[1])), pack2(b0.x, b0.y)),
                                  add2(pack2(__uint_as_float(r[2]), __uint_as_float(r[3])), pack2(b0.z, b0.w)),
                                  add2(pack2(__uint_as_float(r[4]), __uint_as_float(r[5])), pack2(b1.x, b1.y)),
                                  add2(pack2(__uint_as_float(r[6]), __uint_as_float(r[7])), pack2(b1.z, b1.w))};
#pragma unroll
                    for (int e = 0; e < 4; e++) unpack2(p[e], v[2 * e], v[2 * e + 1]);
#else
                    v[0] = __uint_as_float(r[0]) + b0.x; v[1] = __uint_as_float(r[1]) + b0.y;
                    v[2] = __uint_as_float(r[2]) + b0.z; v[3] = __uint_as_float(r[3]) + b0.w;
                    v[4] = __uint_as_float(r[4]) + b1.x; v[5] = __uint_as_float(r[5]) + b1.y;
                    v[6] = __uint_as_float(r[6]) + b1.z; v[7] = __uint_as_float(r[7]) + b1.w;
#endif
                }
                if (!(kDebugFlags(P) & 4)) {
#if LB2_PACKED_F32
#pragma unroll
                    for (int e = 0; e < 4; e++) elu_fast2(v[2 * e], v[2 * e + 1]);
#else
#pragma unroll
                    for (int e = 0; e < 8; e++) v[e] = elu_fast(v[e]);
#endif
                }
            };
            // ordinary layer: pack to fp16 and store one 16-byte chunk row. Padding rows/columns of
            // the row space are written as zeros (whole sectors: partial-sector writes cost L2 fills);
            // layer 1 re-addresses S=21 rows into the S=20 space and must skip them instead.
            auto store_unit = [&](const uint32_t (&r)[8], int u, const float4& b0, const float4& b1) {
                const int h = unit_half(u);
                const int cc = unit_col(u);
                const bool valid = h ? valid2[1] : valid2[0], skip = h ? skip2[1] : skip2[0];
                const int out_row = h ? out_row2[1] : out_row2[0];
                float v[8];
                activate_b(r, unit_addr2(u), b0, b1, v);
                if (!skip) {
                    uint32_t pk[4], pl[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        __half2 hh = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
                        pk[e] = valid ? *reinterpret_cast<uint32_t*>(&hh) : 0u;
                        if (kLo16 && out_mode == kOutLo16) {   // what the fp16 rounding dropped, as a second fp16
                            const float2 hf = __half22float2(hh);
                            __half2 ll = __floats2half2_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
                            pl[e] = valid ? *reinterpret_cast<uint32_t*>(&ll) : 0u;
                        }
                    }
                    const int c8 = (col0 + cc) >> 3;
                    if (kLo16 && out_mode == kOutLo16) {
                        __half* lo = out + (size_t)row_off(lo_chunks + c8, out_row) * 8;
                        if (LB2_L2_HINTS) st_global_v4_hint(lo, make_uint4(pl[0], pl[1], pl[2], pl[3]), st_policy);
                        else *reinterpret_cast<uint4*>(lo) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                    }
                    if (LB2_L2_HINTS)
                        st_global_v4_hint(out + (size_t)row_off(c8, out_row) * 8, make_uint4(pk[0], pk[1], pk[2], pk[3]), st_policy);
                    else
                        *reinterpret_cast<uint4*>(out + (size_t)row_off(c8, out_row) * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
            };
            // Lite mode (kOutFp8): 16 channels (column blocks cc, cc + 8) of one row half. Besides the two fp16 chunk rows
            // the consumer's correction MMAs need, per 16 channels, one 16-byte row of e4m3(a) and one of e4m3((a - fp16(a)) * 2^12):
            // chunk planes lo_chunks + 2 * (channel / 16) and the one behind it.
            auto store_pair_fp8 = [&](const uint32_t (&r)[2][8], int h, int cc) {
                const bool valid = h ? valid2[1] : valid2[0], skip = h ? skip2[1] : skip2[0];
                const int out_row = h ? out_row2[1] : out_row2[0];
                uint32_t qa[4], ql[4];
#pragma unroll
                for (int g = 0; g < 2; g++) {
                    const float* bp = bs + col0 + cc + 8 * g;
                    const float4 b0 = *reinterpret_cast<const float4*>(bp), b1 = *reinterpret_cast<const float4*>(bp + 4);
                    float v[8];
                    activate_b(r[g], tbase + h * 128 + cc + 8 * g, b0, b1, v);
                    uint32_t pk[4], q8[4], q9[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        __half2 hh = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
                        pk[e] = valid ? *reinterpret_cast<uint32_t*>(&hh) : 0u;
                        const float2 hf = __half22float2(hh);
                        q8[e] = cvt_e4m3x2(v[2 * e], v[2 * e + 1]);
#if LB2_PACKED_F32
                        float d0, d1;   // (v - hi) * 2^12, exact
                        unpack2(mul2(add2(pack2(v[2 * e], v[2 * e + 1]), pack2(-hf.x, -hf.y)), pack2(kLoScale, kLoScale)), d0, d1);
                        q9[e] = cvt_e4m3x2(d0, d1);
#else
                        q9[e] = cvt_e4m3x2((v[2 * e] - hf.x) * kLoScale, (v[2 * e + 1] - hf.y) * kLoScale);
#endif
                    }
#ifdef LB2_DEBUG_KNOBS
                    if ((kDebugFlags(P) & 128) && blockIdx.x == 0 && warp == 2 && lane == 5 && it < 40 && g == 0)
                        printf("it %u job %d store: v %g %g a8 %04x lo8 %04x valid %d row %d lo_chunks %d\n", it, jj, v[0], v[1], q8[0], q9[0], (int)valid, out_row, lo_chunks);
#endif
                    qa[2 * g] = valid ? (q8[0] | (q8[1] << 16)) : 0u; qa[2 * g + 1] = valid ? (q8[2] | (q8[3] << 16)) : 0u;
                    ql[2 * g] = valid ? (q9[0] | (q9[1] << 16)) : 0u; ql[2 * g + 1] = valid ? (q9[2] | (q9[3] << 16)) : 0u;
                    if (!skip) {
                        __half* ph = out + (size_t)row_off(((col0 + cc) >> 3) + g, out_row) * 8;
                        if (LB2_L2_HINTS) st_global_v4_hint(ph, make_uint4(pk[0], pk[1], pk[2], pk[3]), st_policy);
                        else *reinterpret_cast<uint4*>(ph) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                }
                if (!skip) {
                    __half* pa = out + (size_t)row_off(lo_chunks + 2 * ((col0 + cc) >> 4), out_row) * 8;
                    __half* pl = pa + (size_t)(uint32_t)chunk_rows * 8;
                    if (LB2_L2_HINTS) {
                        st_global_v4_hint(pa, make_uint4(qa[0], qa[1], qa[2], qa[3]), st_policy);
                        st_global_v4_hint(pl, make_uint4(ql[0], ql[1], ql[2], ql[3]), st_policy);
                    } else {
                        *reinterpret_cast<uint4*>(pa) = make_uint4(qa[0], qa[1], qa[2], qa[3]);
                        *reinterpret_cast<uint4*>(pl) = make_uint4(ql[0], ql[1], ql[2], ql[3]);
                    }
                }
            };

            if (warp == 2 && lane == 0) LB2_TRACE(it, 8);
            mbar_wait(tfull_bar + acc, acc_phase);
            tc_fence_after_sync();
            if (warp == 2 && lane == 0) LB2_TRACE(it, 9);
            if (head) {
                // last trunk layer: the net's final 3x3 conv to ONE channel is folded in here as nine
                // per-tap dot products over this warp's channels, z[t] += sum_c w[t][c] * v[c] (fp32,
                // unrounded v); the heads kernel gathers them. zbuf[part][t][row]; padding rows give 0.
                // Both row halves of a column block go through the taps together: one read of the eight bias values and
                // of the 72 head weights (shared memory: the tensor core's operand port) serves two rows; the next block's
                // TMEM loads are in flight meanwhile. Per row the sums run in the same order as ever.
                if (n_units) {
                    float z0[9], z1[9];
#pragma unroll
                    for (int t = 0; t < 9; t++) z0[t] = z1[t] = 0.0f;
                    uint32_t ra[2][8], rb[2][8];
                    tmem_ld_32x8(unit_addr(0), ra[0]);
                    tmem_ld_32x8(unit_addr(upc), ra[1]);
                    tmem_ld_wait();
                    auto taps = [&](const uint32_t (&r)[2][8], int k) {
                        const float* bp = bs + col0 + k * 8;
                        const float4 b0 = *reinterpret_cast<const float4*>(bp), b1 = *reinterpret_cast<const float4*>(bp + 4);
                        float v0[8], v1[8];
                        activate_b(r[0], unit_addr(k), b0, b1, v0);
                        activate_b(r[1], unit_addr(upc + k), b0, b1, v1);
#pragma unroll
                        for (int t = 0; t < 9; t++) {
                            const float* wt = headw_s + t * 128 + col0 + k * 8;
                            const float4 w0 = *reinterpret_cast<const float4*>(wt), w1 = *reinterpret_cast<const float4*>(wt + 4);
                            z0[t] = fmaf(v0[0], w0.x, z0[t]); z0[t] = fmaf(v0[1], w0.y, z0[t]);
                            z0[t] = fmaf(v0[2], w0.z, z0[t]); z0[t] = fmaf(v0[3], w0.w, z0[t]);
                            z0[t] = fmaf(v0[4], w1.x, z0[t]); z0[t] = fmaf(v0[5], w1.y, z0[t]);
                            z0[t] = fmaf(v0[6], w1.z, z0[t]); z0[t] = fmaf(v0[7], w1.w, z0[t]);
                            z1[t] = fmaf(v1[0], w0.x, z1[t]); z1[t] = fmaf(v1[1], w0.y, z1[t]);
                            z1[t] = fmaf(v1[2], w0.z, z1[t]); z1[t] = fmaf(v1[3], w0.w, z1[t]);
                            z1[t] = fmaf(v1[4], w1.x, z1[t]); z1[t] = fmaf(v1[5], w1.y, z1[t]);
                            z1[t] = fmaf(v1[6], w1.z, z1[t]); z1[t] = fmaf(v1[7], w1.w, z1[t]);
                        }
                    };
#pragma unroll 1
                    for (int k = 0; k < upc; k += 2) {   // upc is even (n_out is a multiple of 32, two column parts)
                        tmem_ld_32x8(unit_addr(k + 1), rb[0]);
                        tmem_ld_32x8(unit_addr(upc + k + 1), rb[1]);
                        taps(ra, k);
                        tmem_ld_wait();
                        if (k + 2 < upc) {
                            tmem_ld_32x8(unit_addr(k + 2), ra[0]);
                            tmem_ld_32x8(unit_addr(upc + k + 2), ra[1]);
                        }
                        taps(rb, k + 1);
                        tmem_ld_wait();
                    }
                    const int zp = (J.zparts + part) * 9;
#pragma unroll
                    for (int t = 0; t < 9; t++) {   // (kept in L2 for the heads kernel like the activations: next to evict_last lines, plain ones go first)
                        if (LB2_L2_HINTS) {
                            st_global_f32_hint(zbuf + row_off(zp + t, out_row2[0]), valid2[0] ? z0[t] : 0.0f, st_policy);
                            st_global_f32_hint(zbuf + row_off(zp + t, out_row2[1]), valid2[1] ? z1[t] : 0.0f, st_policy);
                        } else {
                            zbuf[row_off(zp + t, out_row2[0])] = valid2[0] ? z0[t] : 0.0f;
                            zbuf[row_off(zp + t, out_row2[1])] = valid2[1] ? z1[t] : 0.0f;
                        }
                    }
                }
            } else if (kFp8 && out_mode == kOutFp8) {
                // lite jobs: pair p = (row half p & 1, 16 channels p >> 1); the next pair's TMEM loads are in flight meanwhile
                // (c_out is a multiple of 32, so every warp has an even number of 8-column blocks)
                const int n_pairs = n_units >> 1;
                auto load_pair = [&](int p, uint32_t (&r)[2][8]) {
                    if (p < n_pairs) {
                        const uint32_t a = tbase + (p & 1) * 128 + (p >> 1) * 16;
                        tmem_ld_32x8(a, r[0]);
                        tmem_ld_32x8(a + 8, r[1]);
                    }
                };
                uint32_t ra[2][8], rb[2][8];
                load_pair(0, ra);
                tmem_ld_wait();
#pragma unroll 1
                for (int p = 0; p < n_pairs; p += 2) {
                    load_pair(p + 1, rb);
                    store_pair_fp8(ra, p & 1, (p >> 1) * 16);
                    tmem_ld_wait();
                    load_pair(p + 2, ra);
                    store_pair_fp8(rb, (p + 1) & 1, ((p + 1) >> 1) * 16);
                    tmem_ld_wait();
                }
            } else if (n_units) {
                auto load_group = [&](int u, uint32_t (&r)[G][8]) {
#pragma unroll
                    for (int g = 0; g < G; g++)
                        if (u + g < n_units) tmem_ld_32x8(unit_addr2(u + g), r[g]);
                };
                auto store_group = [&](int u, const uint32_t (&r)[G][8]) {
                    float4 b0, b1;
#pragma unroll
                    for (int g = 0; g < G; g++)
                        if (u + g < n_units) {
                            if (!LB2_EPI_PAIR_HALVES || !(g & 1)) {   // (u is even: G is)
                                const float* bp = bs + col0 + unit_col(u + g);
                                b0 = *reinterpret_cast<const float4*>(bp); b1 = *reinterpret_cast<const float4*>(bp + 4);
                            }
                            store_unit(r[g], u + g, b0, b1);
                        }
                };
                static_assert(!LB2_EPI_PAIR_HALVES || G % 2 == 0, "paired halves need an even group");
#if LB2_EPI_PIPE
                uint32_t ra[G][8], rb[G][8];
                load_group(0, ra);
                tmem_ld_wait();
#pragma unroll 1
                for (int u = 0; u < n_units; u += 2 * G) {
                    load_group(u + G, rb);
                    store_group(u, ra);
                    tmem_ld_wait();
                    load_group(u + 2 * G, ra);
                    store_group(u + G, rb);
                    tmem_ld_wait();
                }
#else
#pragma unroll 1
                for (int u = 0; u < n_units; u += G) {
                    uint32_t ra[G][8];
                    load_group(u, ra);
                    tmem_ld_wait();
                    store_group(u, ra);
                }
#endif
            }
            // accumulator drained: hand it back to the MMA warp (of the leader CTA in pair mode)
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) {
                if (kPair && !leader) { if (LB2_RELAXED_HANDBACK) mbar_arrive_remote_relaxed(tempty_bar + acc, 0); else mbar_arrive_remote(tempty_bar + acc, 0); }
                else mbar_arrive(tempty_bar + acc);
            }
            if (warp == 2 && lane == 0) LB2_TRACE(it, 10);
            // hand the tile to the publisher warp: this warp's stores happen-before its arrive
            if (lane == 0) {
                while (it >= *pub_done + kPubDepth) __nanosleep(32);  // ring slot free? (normally yes)
                mbar_arrive(pub_bar + (it % kPubDepth));
            }
        }
    } else if (warp == 2 + kEpilogueWarps) {
        // ================================ tile publisher ==============================
        // Keeps the gpu-scope release (which waits for the tile's stores to land in L2) off the
        // epilogue's critical path. One lane: wait until all 8 epilogue warps stored tile `it`,
        // then release its flag for the consumers' acquire (cumulative over the mbarrier sync).
        // Its progress counter also bounds how far ahead the scout hands out items.
        // Before the release it drops dead activation tiles from L2 (discard.global.L2: no write-back): input tile c of this
        // layer is dead once the three tiles that read it (c - 1, c, c + 1 of this layer) are complete, and nothing writes its
        // rows before the next layer's tile c, which waits for exactly those three flags. Whoever completes the last of the
        // three — judged before publishing its own flag — discards: so every discard happens-before the flag release that
        // lets the rows be overwritten, and two neighbours finishing together at worst both leave the tile alone. Without
        // this most of a layer's output was written back to HBM after its last read (389 MB per launch at batch 256).
        uint32_t it = 0;
        for (int jj, idx; item_get<true>(item_ring, it, jj, idx); it++) {
            if (lane == 0) mbar_wait(pub_bar + (it % kPubDepth), (it / kPubDepth) & 1);
            __syncwarp();
            if (lane == 0) LB2_TRACE(it, 11);
            if (P.use_flags) {
                const LayerJob& J = jobs[jj];
                const int T = tile_of(idx);
                if (LB2_DISCARD_DEAD && J.in_planes) {
                    const int n_tiles = kPair ? 2 * J.n_items : J.n_items;
                    // lanes 0..4 look at tiles T - 2 .. T + 2. T itself counts as complete, and so does the peer CTA's tile of the
                    // same item (T ^ 1): the item's MMAs — the last readers of its input — were complete before either epilogue ran.
                    const int t = T - 2 + lane;
                    const bool done = lane < 5 && (t == T || (kPair && t == (T ^ 1)) || t < 0 || t >= n_tiles || ld_acquire_gpu(J.flags + t) == epoch);
                    const uint32_t mask = __ballot_sync(0xffffffffu, done);   // bit i: tile T - 2 + i complete (or outside)
                    // this CTA looks after its own input tile and the neighbour on the far side of its peer
                    const int c_lo = kPair ? (rank == 0 ? T - 1 : T) : T - 1, c_hi = kPair ? (rank == 0 ? T : T + 1) : T + 1;
                    for (int c = c_lo; c <= c_hi; c++) {
                        if (c < 0 || c >= n_tiles) continue;
                        const int b = c - (T - 2);   // bit of tile c; its readers are bits b - 1, b, b + 1
                        if (((mask >> (b - 1)) & 7u) != 7u) continue;
                        // 256 rows x 16 bytes = 32 lines of 128 bytes per chunk plane
                        const char* base = reinterpret_cast<const char*>(J.in_base) + (size_t)c * kTileRows * 16 + lane * 128;
                        for (int p = 0; p < J.in_planes; p++) discard_l2_128(base + (size_t)p * J.in_chunk_rows * 16);
                    }
                    __syncwarp();
                }
                if (lane == 0) {
                    fence_proxy_async();  // generic-proxy stores -> visible to other CTAs' TMA loads
                    st_release_gpu(J.flags + T, epoch);
                }
            }
            if (lane == 0) {
                LB2_TRACE(it, 12);
                *pub_done = it + 1;
            }
        }
    } else if (warp == 3 + kEpilogueWarps) {
        // ================================ scout ========================================
        // Leader: hands out the cluster's items (see item_pack) a few items ahead of what has been
        // published. Every CTA: polls (acquire, gpu scope, lanes in parallel) the <= 3 tile flags of
        // the previous layer that cover an item's rows +- halo ahead of the producer and publishes
        // its progress in shared memory — the L2 round trips of the polling stay off the load path.
        const bool dynamic = P.dynamic != 0;
        const int first = P.item_begin + (kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x);
        const int step = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
        int j = 0;
        // kRes: the net this cluster currently draws items from, and which nets have run dry
        int cur_net = kRes ? ((int)(blockIdx.x >> 1) < P.policy_clusters ? 0 : 1) : 0;
        bool dry[2] = {false, false};
        for (uint32_t it = 0;; it++) {
            int jj = (int)kEndJob, idx = 0;
            if (leader) {
                int q = 0, q_end = P.item_end;
                if (lane == 0) {
                    while (it >= *pub_done + kClaimAhead) __nanosleep(20);
                    if (kRes) {
                        // one end marker is drawn from EACH net's counter by every cluster: items + clusters claims per net per launch
                        for (;;) {
                            q = P.net_item_begin[cur_net] + (int)atomicAdd(P.sched + 1 + cur_net, 1u);
                            q_end = P.net_item_end[cur_net];
                            if (q < q_end) break;
                            dry[cur_net] = true;
                            if (dry[cur_net ^ 1]) break;
                            cur_net ^= 1;
                            j = -1;   // the other net's items lie elsewhere in the list: restart the round cursor
                        }
                    } else {
                        q = dynamic ? P.item_begin + (int)(atomicAdd(P.sched, 1u) - P.claim_base) : first + (int)it * step;
                    }
                }
                q = __shfl_sync(0xffffffffu, q, 0);
                q_end = __shfl_sync(0xffffffffu, q_end, 0);
                if (kRes) { if (__shfl_sync(0xffffffffu, j, 0) < 0) j = 0; }
                if (q < q_end) locate_item(P, jobs, q, j, jj, idx);
                if (lane == 0) {
                    st_volatile_shared(geom_ring + it % kClaimRing, geom_pack(it, jj == (int)kEndJob ? nullptr : &jobs[jj]));
                    const uint32_t e = item_pack(it, (uint32_t)jj, (uint32_t)idx);
                    if (kPair) st_remote_shared(item_ring + it % kClaimRing, 1, e);
                    st_volatile_shared(item_ring + it % kClaimRing, e);
                }
            } else {
                item_get<true>(item_ring, it, jj, idx);
            }
            if (jj == (int)kEndJob) break;
            const LayerJob& J = jobs[jj];
            if (P.use_flags && J.dep_job >= 0) {
                int lo, hi;
                dependency_range(J, tile_of(idx), lo, hi);
                for (int sp = 0; sp < J.dep_n_split; sp++) {   // every column split of the producing layer
                    const uint32_t* flags = jobs[J.dep_job + sp].flags;
                    if (lo + lane <= hi)
                        while (ld_acquire_gpu(flags + lo + lane) != epoch) __nanosleep(20);
                }
            }
            __syncwarp();
            if (lane == 0) {
                if (LB2_CONSUMER_PROXY_FENCE == 2) fence_proxy_async();   // flags acquired (all lanes, above) -> fence -> release to the producer
                st_release_cta_shared(deps_ready, it + 1);
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (P.trace && threadIdx.x == 0) {
        unsigned long long* t = P.trace + ((size_t)blockIdx.x * kTraceItems + (kTraceItems - 1)) * kTraceEvents;
        t[2] = global_ns(); t[3] = (unsigned long long)clock64();
    }
    if (kPair) cluster_sync_all();  // nobody exits (or frees TMEM) while its peer may still signal it
    if (warp == 1) { if (kPair) tmem_dealloc_pair<512>(tmem_base); else tmem_dealloc<512>(tmem_base); }
}

// ------------------------------------------------------------------------------------------
// heads. The 3x3 conv to one channel was folded into the last trunk layer's epilogue as per-tap
// partial sums z[half][t][row]; what is left is a 9-point gather per board point.
// ------------------------------------------------------------------------------------------
template <int kParts>
__device__ __forceinline__ float head_gather_parts(const float* __restrict__ zbuf, int chunk_rows, int base_row, int y, int x) {
    float acc = 0.0f;
#pragma unroll
    for (int t = 0; t < 9; t++) {
        const int row = base_row + (y + t / 3 - 1) * 20 + (x + t % 3 - 1);
        if (row >= 0) {  // rows above the first position are implicit zero padding
#pragma unroll
            for (int p = 0; p < kParts; p++) acc += zbuf[(size_t)(p * 9 + t) * chunk_rows + row];
        }
    }
    return acc;
}
// n_parts = column splits of the last trunk layer x kColParts; the loads of one point are independent
// and must be issued together, hence the compile-time part counts
__device__ __forceinline__ float head_gather(const float* __restrict__ zbuf, int chunk_rows, int n_parts, int base_row, int y, int x) {
    return n_parts == kColParts ? head_gather_parts<kColParts>(zbuf, chunk_rows, base_row, y, x)
                                : head_gather_parts<kColParts * kMaxSplit>(zbuf, chunk_rows, base_row, y, x);
}

__device__ __forceinline__ float warp_max(float v) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Both heads in ONE launch of kHeadThreads-thread blocks, all of them resident at once at batch 256 (round 1 ran the value
// head as 64 blocks of 1024 threads that each streamed the whole 370 KB ip matrix through 218 KB of shared memory, and the
// policy head as one such block per position: 320 blocks, one per SM, 2.2 waves, 19 us):
//   blocks [0, value_blocks): value head, block = (group of kVGroup positions, slice of kVSliceRows rows of the 361 x H matrix)
//   the rest:                 policy head, kPolicyPerBlock positions per block, 128 threads each
constexpr int kHeadThreads = 512;
constexpr int kVGroup = 16;        // positions per value block
constexpr int kVSliceRows = 41;    // rows (board points) of the 361 x H inner-product matrix per value block
constexpr int kVSlices = (kPoints + kVSliceRows - 1) / kVSliceRows;   // 9
constexpr int kVHiddenMax = 256;
constexpr int kPolicyPerBlock = kHeadThreads / 128;

// policy head: logit = ELU(b + conv), softmax with temperature over the 361 points (Network.cpp:450-469), un-rotate
// (Network.cpp:820-823). 128 threads per position, three points each.
__device__ __forceinline__ void policy_head_body(const float* __restrict__ zbuf, int chunk_rows, int n_parts,
                                                 const float* __restrict__ bias, const uint8_t* __restrict__ rotation, int ensemble,
                                                 float temp, float* __restrict__ probs, int n, int block, float* smem_f) {
    const int sub = threadIdx.x >> 7, t = threadIdx.x & 127, w = t >> 5;
    const int pos = block * kPolicyPerBlock + sub;
    float* sm = smem_f + sub * 384;                       // [361] this position's probabilities
    float* red = smem_f + kPolicyPerBlock * 384 + sub * 8;  // [4] per-warp partial results
    if (LB2_PDL) grid_dep_wait();
    if (pos >= n) return;   // (all 128 threads of the position: its named barrier is not used)
    float logit[3], e[3];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int p = t + 128 * k;
        logit[k] = -INFINITY;
        if (p < kPoints) {
            const int y = p / kBoard, x = p - y * kBoard;
            logit[k] = elu1(bias[0] + head_gather(zbuf, chunk_rows, n_parts, pos * 400, y, x));
        }
        m = fmaxf(m, logit[k]);
    }
    m = warp_max(m);
    if ((t & 31) == 0) red[w] = m;
    named_bar_sync(1 + sub, 128);
    m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    named_bar_sync(1 + sub, 128);
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        e[k] = (t + 128 * k < kPoints) ? expf(logit[k] / temp - m / temp) : 0.0f;
        s += e[k];
    }
    s = warp_sum(s);
    if ((t & 31) == 0) red[w] = s;
    named_bar_sync(1 + sub, 128);
    s = (red[0] + red[1]) + (red[2] + red[3]);
#pragma unroll
    for (int k = 0; k < 3; k++)
        if (t + 128 * k < kPoints) sm[t + 128 * k] = e[k] / s;
    named_bar_sync(1 + sub, 128);
    const int rot = ensemble ? (pos & 7) : (rotation[pos] & 7);
#pragma unroll
    for (int k = 0; k < 3; k++)
        if (t + 128 * k < kPoints) probs[(size_t)pos * kPoints + t + 128 * k] = sm[rev_rotate_idx(t + 128 * k, rot)];
}

// value head: v = ELU(b + conv) [361]; h = ELU(W1 v + b1); out = (1 + tanh(w2.h + b2)) / 2 (Network.cpp:731-737, 395-423).
// Block (g, s) multiplies rows [s * 41, s * 41 + 41) of W1 (one bulk copy into shared memory, <= 42 KB) with the matching
// points of positions [16 g, 16 g + 16) and leaves its partial sums in `partial`; the block that arrives last at the
// group's counter adds the nine slices IN SLICE ORDER (so a position's result does not depend on which block that was, nor
// on the batch around it), applies b1 + ELU, the second inner product and the tanh, and zeroes the counter for the next launch.
__device__ __forceinline__ void value_head_body(const HeadArgs& A, int block, uint8_t* vsm) {
    const int hidden = A.hidden, n = A.n_value;
    float* w_s = reinterpret_cast<float*>(vsm);                                  // [41][hidden]; later h [16][hidden]
    const int w_floats = (kVSliceRows * hidden > kVGroup * kVHiddenMax) ? kVSliceRows * hidden : kVGroup * kVHiddenMax;
    float* v_s = w_s + w_floats;                                                 // [41][16]
    uint64_t* full = reinterpret_cast<uint64_t*>(v_s + kVSliceRows * kVGroup);
    uint32_t* last_s = reinterpret_cast<uint32_t*>(full + 1);
    const int tid = threadIdx.x;
    const int g = block / kVSlices, s = block - g * kVSlices;
    const int row0 = s * kVSliceRows, rows = min(kVSliceRows, kPoints - row0);
    const int pos0 = g * kVGroup;
    if (tid == 0) {
        mbar_init(full, 1);
        fence_mbar_init();
        fence_proxy_async_smem();
        const uint32_t bytes = (uint32_t)(rows * hidden * sizeof(float));
        mbar_arrive_expect_tx(full, bytes);
        bulk_load_1d(w_s, A.ip1_wt + (size_t)row0 * hidden, bytes, full);
    }
    if (LB2_PDL) grid_dep_wait();   // the weights above do not depend on the trunk; zbuf does
    for (int i = tid; i < kVGroup * rows; i += kHeadThreads) {   // consecutive threads: consecutive points of one position
        const int gp = i / rows, r = i - gp * rows;
        float v = 0.0f;
        if (pos0 + gp < n) {
            const int p = row0 + r, y = p / kBoard, x = p - y * kBoard;
            v = elu1(A.v_bias[0] + head_gather(A.v_zbuf, A.v_chunk_rows, A.v_parts, (pos0 + gp) * 400, y, x));
        }
        v_s[r * kVGroup + gp] = v;
    }
    __syncthreads();
    mbar_wait(full, 0);
    const int o = tid & (kVHiddenMax - 1), half = tid >> 8;   // output o, positions [8 half, 8 half + 8) of the group
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; j++) a[j] = 0.0f;
    if (o < hidden) {
#pragma unroll 4
        for (int r = 0; r < rows; r++) {
            const float wv = w_s[r * hidden + o];
            const float4 v0 = *reinterpret_cast<const float4*>(v_s + r * kVGroup + half * 8);
            const float4 v1 = *reinterpret_cast<const float4*>(v_s + r * kVGroup + half * 8 + 4);
            a[0] = fmaf(wv, v0.x, a[0]); a[1] = fmaf(wv, v0.y, a[1]); a[2] = fmaf(wv, v0.z, a[2]); a[3] = fmaf(wv, v0.w, a[3]);
            a[4] = fmaf(wv, v1.x, a[4]); a[5] = fmaf(wv, v1.y, a[5]); a[6] = fmaf(wv, v1.z, a[6]); a[7] = fmaf(wv, v1.w, a[7]);
        }
        float* part = A.v_partial + ((size_t)(g * kVSlices + s) * kVGroup + half * 8) * kVHiddenMax + o;
#pragma unroll
        for (int j = 0; j < 8; j++) part[(size_t)j * kVHiddenMax] = a[j];
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) *last_s = (atomicAdd(A.v_count + g, 1u) == (uint32_t)(kVSlices - 1)) ? 1u : 0u;
    __syncthreads();
    if (!*last_s) return;
    __threadfence();
    float* h_s = w_s;   // (every thread of this block is past its reads of the weight slice)
    if (o < hidden) {
        const float* part = A.v_partial + ((size_t)g * kVSlices * kVGroup + half * 8) * kVHiddenMax + o;
        const float b1 = A.ip1_b[o];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            float sum = __ldcg(part + (size_t)j * kVHiddenMax);
            for (int s2 = 1; s2 < kVSlices; s2++) sum += __ldcg(part + ((size_t)s2 * kVGroup + j) * kVHiddenMax);
            h_s[(half * 8 + j) * hidden + o] = elu1(sum + b1);
        }
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;   // 16 warps: one position each
    if (pos0 + warp < n) {
        float d = 0.0f;
        for (int k = lane; k < hidden; k += 32) d = fmaf(A.ip2_w[k], h_s[warp * hidden + k], d);
        d = warp_sum(d);
        if (lane == 0) A.winrate[pos0 + warp] = (1.0f + tanhf(d + A.ip2_b[0])) * 0.5f;
    }
    if (tid == 0) A.v_count[g] = 0;
}

__global__ void __launch_bounds__(kHeadThreads) heads_kernel(const HeadArgs A) {
    extern __shared__ __align__(128) uint8_t hsm[];
    const int value_blocks = ((A.n_value + kVGroup - 1) / kVGroup) * kVSlices;
    unsigned long long* tr = (A.trace && threadIdx.x == 0 && (int)blockIdx.x < A.trace_ctas)
                                 ? A.trace + ((size_t)blockIdx.x * kTraceItems + (kTraceItems - 2)) * kTraceEvents : nullptr;
    if (tr) tr[0] = global_ns();
    if ((int)blockIdx.x < value_blocks)
        value_head_body(A, blockIdx.x, hsm);
    else
        policy_head_body(A.p_zbuf, A.p_chunk_rows, A.p_parts, A.p_bias, A.rotation, A.ensemble, A.temp, A.probs, A.n_policy,
                         blockIdx.x - value_blocks, reinterpret_cast<float*>(hsm));
    if (tr) tr[8] = global_ns();
}

// AVERAGE_ALL: one thread per output element; the 8 addends are summed in the reference's order.
__global__ void __launch_bounds__(256) ensemble_mean_kernel(const MeanArgs A) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int np = A.n_policy * kPoints;
    if (t < np) {
        const int i = t / kPoints, idx = t - i * kPoints;
        const float* src = A.probs8 + (size_t)i * 8 * kPoints + idx;
        float s = src[0];
#pragma unroll
        for (int r = 1; r < 8; r++) s += src[(size_t)r * kPoints];
        A.probs[t] = s / 8.0f;
    } else if (t - np < A.n_value) {
        const int i = t - np;
        float s = A.win8[8 * i];
#pragma unroll
        for (int r = 1; r < 8; r++) s += A.win8[8 * i + r];
        A.win[i] = s / 8.0f;
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
// host-side addresses of the kernels whose arguments hold caller pointers (the host API patches their graph nodes)
const void* kernel_address(int which) {
    return which == 0 ? (const void*)expand_planes_kernel : (which == 1 ? (const void*)heads_kernel : (const void*)ensemble_mean_kernel);
}

cudaError_t launch_ensemble_mean(const MeanArgs& a, cudaStream_t st) {
    const int total = a.n_policy * kPoints + a.n_value;
    if (total == 0) return cudaSuccess;
    ensemble_mean_kernel<<<(total + 255) / 256, 256, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_expand(const ExpandArgs& a, cudaStream_t st) {
    const size_t threads = (size_t)a.n * 441 * a.n_nets + (a.pf_bytes + 127) / 128;
    expand_planes_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(a);
    return cudaGetLastError();
}

// Instances: plain-only epilogue (kOutModes 0) and one that can also store the split-operand planes (3), each as single CTA,
// CTA pair, CTA pair with resident weights.
#define LB2_FOR_EACH_TRUNK(X) X(false, false, 0) X(true, false, 0) X(true, true, 0) X(false, false, 1) X(true, false, 1) X(true, true, 1) X(false, false, 2) X(true, false, 2) X(true, true, 2)
cudaError_t trunk_kernel_setup() {
    cudaError_t e;
#define LB2_SET(p, r, m)                                                                                                       \
    if ((e = cudaFuncSetAttribute(trunk_kernel<p, r, m>, cudaFuncAttributeMaxDynamicSharedMemorySize,                          \
                                  r ? kTrunkSmemBytesRes : kTrunkSmemBytes)) != cudaSuccess) return e;
    LB2_FOR_EACH_TRUNK(LB2_SET)
#undef LB2_SET
    return cudaFuncSetAttribute(heads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
}

cudaError_t launch_trunk(const TrunkParams& p, int grid, bool cooperative, bool pair, bool resident, int out_modes, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kTrunkThreads);
    cfg.dynamicSmemBytes = resident ? kTrunkSmemBytesRes : kTrunkSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[3];
    int na = 0;
    if (LB2_PDL && pair) {   // (the single-CTA form is a cooperative launch, which excludes it)
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        na++;
    }
    if (pair) {
        // CTA pairs: clusters of 2 are always co-scheduled; with one CTA per SM and grid <= #SMs
        // every cluster is resident, which is all the dataflow needs
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        na++;
    } else {
        attr[na].id = cudaLaunchAttributeCooperative;
        attr[na].val.cooperative = cooperative ? 1 : 0;
        na++;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    resident = resident && pair;
#define LB2_LAUNCH(p_, r_, m_) \
    if (pair == p_ && resident == r_ && out_modes == m_) return cudaLaunchKernelEx(&cfg, trunk_kernel<p_, r_, m_>, p);
    LB2_FOR_EACH_TRUNK(LB2_LAUNCH)
#undef LB2_LAUNCH
    return cudaErrorInvalidValue;
}

size_t heads_partial_floats(int n_value) { return (size_t)((n_value + kVGroup - 1) / kVGroup) * kVSlices * kVGroup * kVHiddenMax; }
size_t heads_count_words(int n_value) { return (size_t)((n_value + kVGroup - 1) / kVGroup); }

cudaError_t launch_heads(const HeadArgs& a, cudaStream_t st) {
    const int value_blocks = ((a.n_value + kVGroup - 1) / kVGroup) * kVSlices;
    const int blocks = value_blocks + (a.n_policy + kPolicyPerBlock - 1) / kPolicyPerBlock;
    if (blocks == 0) return cudaSuccess;
    size_t smem = (size_t)kPolicyPerBlock * (384 + 8) * sizeof(float);  // policy head scratch
    if (a.n_value) {
        const size_t w_floats = std::max(kVSliceRows * a.hidden, kVGroup * kVHiddenMax);
        smem = std::max(smem, (w_floats + kVSliceRows * kVGroup) * sizeof(float) + 64);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(kHeadThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = LB2_PDL ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, heads_kernel, a);
}

}  // namespace lb2
