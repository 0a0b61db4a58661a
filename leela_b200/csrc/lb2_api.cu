// lb2_api.cu — host side of the C ABI declared in include/leela_b200.h.
// Owns devices, replicated weights, per-device workspaces, TMA tensor maps and the batch
// pipeline: pinned staging -> H2D -> expand -> trunk (tcgen05) -> heads -> D2H.
//
// Host pipeline in one paragraph: every device has one compute stream (the activation workspace is shared, so kernels
// of different calls run one after the other there) and kIoSlots I/O slots (own copy stream, device and pinned
// buffers). A host-buffer call — or one device batch of the submit queue — owns a slot: H2D on the slot's stream,
// the kernels as ONE CUDA-graph launch on the compute stream (graphs are cached per batch shape; the kernels keep
// their scheduling state on the device, so a launch sequence is replayable), D2H on the slot's stream. Calls larger
// than max_batch are cut into chunks that are dealt to (device, slot) pairs as they become free — whole batches per
// device, never slivers of one batch on every device.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/leela_b200.h"
#include "lb2_kernels.cuh"

#ifndef LB2_RESIDENT_DEFAULT
#define LB2_RESIDENT_DEFAULT 2   // 0 off, 1 whenever the layers fit, 2 when both nets run in the launch. Measured at batch 256 once the
                                 // remote arrives had lost their fences: both nets 324 -> 311 us (the clusters of a net see only its own
                                 // item mix: no short value item waits for the accumulator behind a policy epilogue), one net alone
                                 // +-1 % (policy) / +4 % (value): hence 2
#endif

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CU_TRY(expr)                                                                                  \
    do {                                                                                              \
        cudaError_t e_ = (expr);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(LB2_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_tiled = nullptr;

int load_driver_entry() {
    if (g_encode_tiled) return LB2_OK;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CU_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(LB2_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
    return LB2_OK;
}

struct HostConv {
    int k, c_in, c_out;
    std::vector<float> w, b;
};
struct HostIp {
    int n_in, n_out;
    std::vector<float> w, b;
};

// One way of cutting a trunk layer into column splits of c_out / n_split channels, every split packed like a layer of its own.
struct SplitPacks {
    int n_split = 1;
    // packed weights [precision: fp16 / lite / full][0 = one CTA per item, 1 = CTA-pair packing][split] (see SlabPacker)
    __half* w[3][2][lb2::kMaxSplit] = {};
    float* bias[lb2::kMaxSplit] = {nullptr, nullptr};
};

// One trunk layer on a device. `nat`: a layer wider than 128 output channels is stored (and run) as two column splits.
// `small`: layers of up to 128 channels packed as two splits as well, for small batches — a pass over a few positions is bound
// by the latency of its chained layers (one 256-row item per layer and net, ~12 us each, while most clusters idle); with every
// layer cut in two, two clusters share an item's MMAs and epilogue. CTA-pair packing only; never the last (fused-head) layer.
struct TrunkLayerDev {
    int k, c_in, c_out;
    SplitPacks nat, small;
    bool has_small = false;
    int q = 0;   // lite mode: the packed weights carry the factor 2^q
};

// One net replicated on one device, with its workspace.
struct NetDev {
    std::vector<TrunkLayerDev> trunk;
    int head_c_in = 0;
    float *head_wt[lb2::kMaxSplit] = {nullptr, nullptr};  // final conv weights per column split of the last trunk layer, [9 taps][c_in / n_split]
    float* head_b = nullptr;
    int hidden = 0;
    float *ip1_wt = nullptr, *ip1_b = nullptr, *ip2_w = nullptr, *ip2_b = nullptr;
    // workspace
    int cap = 0;
    int width = 0;           // widest trunk c_out
    int rows5 = 0, rows3 = 0;  // chunk-plane rows of the S=21 / S=20 buffers
    uint32_t* planes = nullptr;     // input of lb2_debug_trunk
    __half* x0 = nullptr;           // first-conv input (S=21 row space)
    __half* act[2] = {nullptr, nullptr};
    float* zbuf = nullptr;          // fused-head partial sums [splits * parts][9][rows3]
    float* head_partial = nullptr;  // value head: partial inner products per (position group, matrix slice), see heads_kernel
    uint32_t* head_count = nullptr; // ... and the groups' arrival counters (zero at rest)
    uint32_t* flags = nullptr;
    int flags_stride = 0;
    CUtensorMap tm_x0, tm_act[2];
};

constexpr int kIoSlots = 2;
constexpr int kDispatchersPerDevice = 4;   // threads of the submit queue per device

// Input/output buffers of one host-buffer call (or one batch of the submit queue) in flight on a device. Two slots per
// device let the copies of one call overlap the kernels of another.
struct IoSlot {
    cudaStream_t stream = nullptr;          // copies of this slot
    cudaEvent_t ev_in = nullptr, ev_done = nullptr, ev_out = nullptr;
    int cap = 0;
    bool busy = false;
    // 0: free; 1: taken, work being enqueued; 2: kernels enqueued and ev_done recorded behind them; 3: taken by a dispatcher of
    // the submit queue that is letting requests accumulate (linger_for_device)
    std::atomic<int> phase{0};
    uint32_t* d_planes[2] = {nullptr, nullptr};
    uint8_t* d_rot = nullptr;
    float *d_probs = nullptr, *d_win = nullptr;
    uint32_t* h_planes[2] = {nullptr, nullptr};  // pinned staging for pageable caller buffers
    uint8_t* h_rot = nullptr;
    float *h_probs = nullptr, *h_win = nullptr;
};

// What one evaluation launches on a device: the shape of the batch plus every pointer the kernels take.
struct EvalKey {
    int n = 0;                  // device positions (8 per input position for an ensemble)
    int limit[2] = {0, 0};      // trunk layers to run per net (0 = net not run; beyond the trunk = whole net with head)
    bool ensemble = false;
    bool profile = false;       // CUDA events around the trunk (graph nodes recording caller-chosen events)
    uint32_t temp_bits = 0;     // softmax temperature
    const void* in[3] = {nullptr, nullptr, nullptr};   // policy planes, value planes, rotation (device)
    void* out[2] = {nullptr, nullptr};                 // probabilities, winrates (device)
    void* mean[2] = {nullptr, nullptr};                // ensemble: where the 8-symmetry means go
    bool same_shape(const EvalKey& o) const {
        return n == o.n && limit[0] == o.limit[0] && limit[1] == o.limit[1] && ensemble == o.ensemble && profile == o.profile &&
               temp_bits == o.temp_bits && (in[0] != nullptr) == (o.in[0] != nullptr) && (in[1] != nullptr) == (o.in[1] != nullptr);
    }
    bool same_pointers(const EvalKey& o) const {
        return !memcmp(in, o.in, sizeof in) && !memcmp(out, o.out, sizeof out) && !memcmp(mean, o.mean, sizeof mean);
    }
};

struct TrunkLaunch {
    lb2::TrunkParams P;
    int grid = 0;
    bool cooperative = false, pair = false, resident = false;
    int modes = 0;
};

// Every kernel argument of one evaluation, ready to launch (directly or as graph nodes).
struct EvalPlan {
    lb2::ExpandArgs ea;
    std::vector<TrunkLaunch> trunk;   // one launch (dataflow mode) or one per round (trunk_mode 0)
    lb2::HeadArgs ha;
    bool heads = false;
    lb2::MeanArgs ma;
    bool mean = false;
    const __half* last_act[2] = {nullptr, nullptr};   // lb2_debug_trunk: the activations of the last layer run
    int kernels = 0;
};

// A cached, instantiated CUDA graph of one evaluation shape.
struct GraphEntry {
    EvalKey key;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    cudaGraphNode_t n_expand = nullptr, n_heads = nullptr, n_mean = nullptr, n_ev0 = nullptr, n_ev1 = nullptr;
    EvalPlan plan;      // the kernel arguments the nodes currently hold
    int tag = 0;        // whose buffers: I/O slot index, or -1 for calls on device pointers
    long last_use = 0;
    int kernels = 0;
};

struct DeviceState {
    int id = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;        // compute stream: every kernel that touches the activation workspace
    NetDev net[2];
    uint8_t* rot = nullptr;               // rotation buffer of lb2_debug_trunk
    int cap = 0;                          // ... and its capacity
    uint32_t* sched = nullptr;            // lb2::TrunkParams::sched
    unsigned long long* trace = nullptr;  // debug timeline buffer (option "trace")
    // option "profile_trunk": (start, stop) event pairs around trunk launches; the events come from a pool that
    // "profile_reserve" fills ahead of time, so that no event is created inside a timed loop
    std::vector<cudaEvent_t> prof_events, prof_pool;
    IoSlot slots[kIoSlots];
    std::mutex mu;                        // guards everything below and the enqueueing on this device
    cudaEvent_t ev_last = nullptr;        // orders an evaluation on one stream behind the previous one on another
    bool last_on_user_stream = false;     // the most recent evaluation ran on a caller's stream (ev_last marks its end)
    std::vector<GraphEntry> graphs;       // cache, least recently used entry replaced
    std::map<std::vector<long>, int> seen;   // how often a shape has been launched directly (a graph is built on the second use)
    long use_counter = 0;
    long graphs_epoch = 0;                // lb2_ctx::option_epoch the cached graphs were built under
};

// A batch of the submit queue: requests of one kind and temperature packed into pinned buffers.
struct QueueRequest {
    float* out;
    int n;
    lb2_callback cb;
    void* user;
};
struct QueueBatch {
    int kind = 0;
    float temp = 1.0f;
    int n = 0;
    uint32_t* planes = nullptr;   // pinned [cap][361]
    uint8_t* rot = nullptr;       // pinned [cap]
    float* out = nullptr;         // pinned [cap][361] (policy) or [cap] (value)
    int cap = 0;
    std::vector<QueueRequest> req;
};

}  // namespace

struct lb2_net {
    lb2_ctx* ctx;
    int kind;
    bool finalized = false;
    std::vector<HostConv> convs;
    std::vector<HostIp> ips;
};

struct lb2_ctx {
    std::vector<std::unique_ptr<DeviceState>> dev;
    std::unique_ptr<lb2_net> nets[2];
    std::string backend;
    std::mutex opt_mu;     // options below (read under it when a call starts)
    long trunk_mode = 1;
    long max_batch = 256;  // batch-256 chunks keep both nets' ping-pong activations L2-resident (measured best)
    long profile_trunk = 0;
    long cta_pair = 1;
    long dynamic_items = 1;
    long use_graphs = 1;   // launch the kernels of an evaluation as one cached CUDA graph (from the second use of a shape on)
    long resident_weights = LB2_RESIDENT_DEFAULT;   // keep each CTA's half of a layer's weights in shared memory across the layer's items
    // Trunk precision per net (kPrecFp16 / kPrecLite / kPrecFull, see SlabPacker). Default: the value net in lite mode — north_star
    // demands the value within 1e-3 of the reference, which fp16 operands alone miss (3.3e-3); the policy net in fp16.
    long precision[2] = {0, 1};
    long policy_clusters = -1;   // resident mode: clusters that prefer the policy net (-1 = split by estimated work)
    long group_positions = 0;     // net-major launches run the batch in groups of this many positions (0 = whole batch at once).
                                  // Measured at batch 256 with groups of 128: DRAM write-back per launch 389 -> 94 MB, L2 hit rate 56 -> 74 %,
                                  // and the launch 5 % SLOWER (634 k -> 705 k cycles: 100 items per layer and group leave the clusters waiting
                                  // at every layer transition; HBM traffic was never the limit) -> off
    long small_batch = 48;        // device passes of up to this many positions run every layer as two column splits (0 = never)
    long queue_linger = 1; // 1: a dispatcher of the submit queue with a free I/O slot lets requests accumulate while the device is still
                           // computing the other slot's batch (see linger_for_device); 0: it takes what is there at once
    long spin_wait = 1;    // 1: a blocking call polls its completion event (yielding the core between polls); 0: it sleeps on it
    std::atomic<long> option_epoch{0};   // bumped by lb2_set_option: cached graphs of older epochs are rebuilt
    std::atomic<long> launches{0}, graph_launches{0};
    std::atomic<long> stat_positions{0}, stat_batches{0}, stat_requests{0};  // async queue: positions, device batches, requests
    std::mutex slot_mu;                 // guards IoSlot::busy of every device
    std::condition_variable slot_cv;
    std::atomic<unsigned> next_dev{0};  // round-robin start of the search for a free slot
    // pinned-range cache: pointers a caller registered (lb2_register_host_buffer) or that were found page-locked before
    std::mutex pin_mu;
    std::vector<std::pair<uintptr_t, uintptr_t>> pinned_ranges;   // [begin, end)
    std::vector<std::pair<uintptr_t, bool>> pin_cache;            // recent cudaPointerGetAttributes answers by address
    std::vector<void*> registered;                                // ranges this library page-locked itself
    // async submission
    std::mutex q_mu;
    std::condition_variable q_cv, q_idle, q_free;
    std::deque<QueueBatch*> q_ready;       // sealed batches, oldest first
    QueueBatch* q_open[2] = {nullptr, nullptr};   // the batch new requests of each kind are appended to
    std::vector<QueueBatch*> q_free_list;
    std::vector<std::unique_ptr<QueueBatch>> q_all;
    int q_cap = 0;                         // positions per batch buffer
    bool worker_run = false;
    int workers_busy = 0;
    std::atomic<long> q_waiting{0};        // positions in the open and sealed batches (read without q_mu by lingering dispatchers)
    int q_idle_workers = 0;                // dispatchers asleep on q_cv (a submitter only signals when there is one)
    std::vector<std::thread> workers;      // kDispatchersPerDevice per device
    std::string q_error;                   // text of the last asynchronous failure (reported by lb2_drain)
};

namespace {

enum { kPrecFp16 = 0, kPrecLite = 1, kPrecFull = 2 };

// --------------------------------------------------------------------------------------------
// weights
// --------------------------------------------------------------------------------------------
void tap_groups(int k, int* n, int* begin, int* end) {
    if (k == 3) { *n = 1; begin[0] = 0; end[0] = 9; }
    else { *n = 3; begin[0] = 0; end[0] = 9; begin[1] = 9; end[1] = 17; begin[2] = 17; end[2] = 25; }
}

// Precision of a net's trunk (options "policy_precision" / "value_precision"): kPrecFp16 / kPrecLite / kPrecFull.
// Virtual K slabs of a layer. A slab is 16 input channels = 32 bytes per (tap, output channel): two 16-byte core-matrix rows.
//   fp16 (kPrecFp16):  slab s = channels [16 s, 16 s + 16) of Wh = fp16(w).
//   full (kPrecFull):  fp32 weight w = Wh + Wl (+ 2^-22 w), Wl = fp16(w - Wh); activation a = hi + lo likewise. The K loop runs
//                      over [hi x Wh | hi x Wl | lo x Wh] (the first layer, whose inputs are 0/1, has no lo term): virtual slab v
//                      takes channels 16 (v % n) from Wh (v < n or v >= 2n) or Wl (n <= v < 2n).
//   lite (kPrecLite):  the weights are scaled by 2^q (q from the layer's largest |w|, so that |w| 2^q < 2^15): Ws = w 2^q,
//                      Wh = fp16(Ws), Wl = Ws - Wh (|Wl| <= 8). Slabs [0, n): Wh, multiplied by hi in a kind::f16 MMA. Slabs
//                      [n, 2n) are e4m3 (kind::f8f6f4, K = 32): row 0 = e4m3(Wl) of the 16 channels, multiplied by e4m3(a);
//                      row 1 = e4m3(w 2^(q - 12)), multiplied by e4m3((a - hi) 2^12). Everything accumulates 2^q x the true
//                      sum in the same fp32 accumulator; the epilogue multiplies by 2^-q. The corrections are 2^-11 of the
//                      main term, so their 4-bit mantissas leave an error of ~2^-15. The first layer (binary inputs) runs
//                      [hi x Wh | hi x Wl] in fp16 like the full mode, scaled.
struct SlabPacker {
    const HostConv& c;
    int mode, q, n_real;
    bool first;
    int n_virtual() const { return mode == kPrecFp16 ? n_real : (mode == kPrecFull ? (first ? 2 : 3) * n_real : 2 * n_real); }
    float scaled(int co, int ci, int t) const { return ldexpf(c.w[((size_t)co * c.c_in + ci) * c.k * c.k + t], mode == kPrecLite ? q : 0); }
    // the 16 bytes of (virtual slab v, tap t, core-matrix row j, output channel co)
    void row(int v, int t, int j, int co, uint8_t* dst) const {
        const int s = v % n_real, term = v / n_real;
        if (mode == kPrecLite && !first && term == 1) {
            int variant = 0;
#ifdef LB2_DEBUG_KNOBS
            if (const char* v = getenv("LB2_LITE_VARIANT")) variant = atoi(v);   // bring-up: 1 zero Wl8, 2 zero W8, 4 swap the rows, 8 swap channel pairs
#endif
            const int jj = (variant & 4) ? 1 - j : j;
            for (int e = 0; e < 16; e++) {
                const float ws = scaled(co, 16 * s + ((variant & 8) ? (e ^ 1) : e), t);
                float val = jj == 0 ? ws - __half2float(__float2half_rn(ws)) : ldexpf(ws, -12);
                if ((variant & 1) && jj == 0) val = 0.0f;
                if ((variant & 2) && jj == 1) val = 0.0f;
                dst[e] = (uint8_t)__nv_cvt_float_to_fp8(val, __NV_SATFINITE, __NV_E4M3);
            }
            return;
        }
        __half* d = reinterpret_cast<__half*>(dst);
        for (int e = 0; e < 8; e++) {
            const float ws = scaled(co, 16 * s + 8 * j + e, t);
            const __half wh = __float2half_rn(ws);
            d[e] = term == 1 ? __float2half_rn(ws - __half2float(wh)) : wh;
        }
    }
};

// [slab][tap group][tap][2 rows][c_out][16 B] — the smem image of each pipeline stage; `pair`: the CTA-pair packing
// [slab][tap group][rank][tap][2 rows][c_out/2][16 B], rank r holding output channels [r c_out/2, (r+1) c_out/2) — the half
// of the MMA's B operand that CTA r of the pair stages.
std::vector<__half> pack_trunk_weights(const HostConv& c, int mode, bool first, int q, bool pair) {
    SlabPacker pk{c, mode, q, c.c_in / 16, first};
    const int kk = c.k * c.k, nv = pk.n_virtual(), ranks = pair ? 2 : 1, nh = c.c_out / ranks;
    std::vector<__half> out((size_t)kk * 16 * nv * c.c_out);
    uint8_t* o = reinterpret_cast<uint8_t*>(out.data());
    int ng, gb[3], ge[3];
    tap_groups(c.k, &ng, gb, ge);
    for (int v = 0; v < nv; v++)
        for (int g = 0; g < ng; g++)
            for (int r = 0; r < ranks; r++)
                for (int t = gb[g]; t < ge[g]; t++)
                    for (int j = 0; j < 2; j++)
                        for (int n = 0; n < nh; n++, o += 16) pk.row(v, t, j, r * nh + n, o);
    return out;
}

// lite mode: q with max |w| 2^q in [2^14, 2^15)
int lite_scale_exponent(const HostConv& c) {
    float m = 0.0f;
    for (float w : c.w) m = std::max(m, std::fabs(w));
    if (!(m > 0.0f)) return 0;
    int e;
    std::frexp(m, &e);   // m = f 2^e, f in [0.5, 1)
    return std::max(-8, std::min(30, 15 - e));
}

template <class T>
int upload(T** dst, const void* src, size_t bytes) {
    CU_TRY(cudaMalloc(reinterpret_cast<void**>(dst), bytes));
    CU_TRY(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
    return LB2_OK;
}

int n_splits(int c_out) { return c_out > 128 ? 2 : 1; }

int upload_net(const lb2_net* net, NetDev* nd) {
    const size_t nconv = net->convs.size();
    nd->trunk.clear();
    nd->width = 0;
    for (size_t l = 0; l + 1 < nconv; l++) {
        const HostConv& c = net->convs[l];
        TrunkLayerDev t;
        t.k = c.k; t.c_in = c.c_in; t.c_out = c.c_out;
        t.q = lite_scale_exponent(c);   // one scale for the whole layer (all column splits)
        const bool first = (l == 0);
        auto pack_splits = [&](SplitPacks& sp_set, int n_split, bool pair_only) -> int {
            sp_set.n_split = n_split;
            const int w = c.c_out / n_split;
            for (int sp = 0; sp < n_split; sp++) {
                HostConv part;   // output channels [sp*w, (sp+1)*w): a contiguous block of the OIHW array
                part.k = c.k; part.c_in = c.c_in; part.c_out = w;
                const size_t per_out = (size_t)c.c_in * c.k * c.k;
                part.w.assign(c.w.begin() + (size_t)sp * w * per_out, c.w.begin() + (size_t)(sp + 1) * w * per_out);
                part.b.assign(c.b.begin() + sp * w, c.b.begin() + (sp + 1) * w);
                int rc = LB2_OK;
                for (int mode = 0; mode < 3; mode++)
                    for (int pair = pair_only ? 1 : 0; pair < 2; pair++) {
                        std::vector<__half> pk = pack_trunk_weights(part, mode, first, t.q, pair != 0);
                        if ((rc = upload(&sp_set.w[mode][pair][sp], pk.data(), pk.size() * sizeof(__half)))) return rc;
                    }
                if ((rc = upload(&sp_set.bias[sp], part.b.data(), part.b.size() * sizeof(float)))) return rc;
            }
            return LB2_OK;
        };
        int rc = pack_splits(t.nat, n_splits(c.c_out), false);
        if (rc) return rc;
        t.has_small = t.nat.n_split == 1 && l + 2 < nconv && (c.c_out / 2) % 32 == 0;
        if (t.has_small && (rc = pack_splits(t.small, 2, true))) return rc;
        nd->trunk.push_back(t);
        nd->width = std::max(nd->width, c.c_out);
    }
    const HostConv& h = net->convs.back();
    nd->head_c_in = h.c_in;
    {
        const int ns = n_splits(h.c_in), w = h.c_in / ns;
        for (int sp = 0; sp < ns; sp++) {
            std::vector<float> hwt((size_t)9 * w);
            for (int c = 0; c < w; c++)
                for (int t = 0; t < 9; t++) hwt[(size_t)t * w + c] = h.w[(size_t)(sp * w + c) * 9 + t];
            int rc = upload(&nd->head_wt[sp], hwt.data(), hwt.size() * sizeof(float));
            if (rc) return rc;
        }
    }
    int rc = upload(&nd->head_b, h.b.data(), sizeof(float));
    if (rc) return rc;
    if (net->kind == LB2_VALUE) {
        const HostIp& a = net->ips[0];
        const HostIp& b = net->ips[1];
        nd->hidden = a.n_out;
        std::vector<float> wt((size_t)a.n_in * a.n_out);  // transpose to [n_in][n_out] for coalesced reads
        for (int o = 0; o < a.n_out; o++)
            for (int i = 0; i < a.n_in; i++) wt[(size_t)i * a.n_out + o] = a.w[(size_t)o * a.n_in + i];
        if ((rc = upload(&nd->ip1_wt, wt.data(), wt.size() * sizeof(float)))) return rc;
        if ((rc = upload(&nd->ip1_b, a.b.data(), a.b.size() * sizeof(float)))) return rc;
        if ((rc = upload(&nd->ip2_w, b.w.data(), b.w.size() * sizeof(float)))) return rc;
        if ((rc = upload(&nd->ip2_b, b.b.data(), sizeof(float)))) return rc;
    }
    return LB2_OK;
}

// --------------------------------------------------------------------------------------------
// workspaces + tensor maps
// --------------------------------------------------------------------------------------------
int round_up(int v, int m) { return (v + m - 1) / m * m; }

// Activation buffer [chunks][rows][8] fp16 viewed as 3-D {64 elems = 8 rows x 8 ch, rows/8, chunks};
// a box {64, (256+2*halo)/8, 2} is one [rows x 16 channels] A slab in core-matrix order.
int make_act_tmap(CUtensorMap* tm, __half* base, int rows, int chunks, int halo) {
    cuuint64_t gdim[3] = {64, (cuuint64_t)rows / 8, (cuuint64_t)chunks};
    cuuint64_t gstride[2] = {128, (cuuint64_t)rows * 16};
    cuuint32_t box[3] = {64, (cuuint32_t)(lb2::kTileRows + 2 * halo) / 8, 2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, base, gdim, gstride, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(LB2_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return LB2_OK;
}

void free_workspace(NetDev* nd) {
    cudaFree(nd->planes); cudaFree(nd->act[0]); cudaFree(nd->act[1]);
    cudaFree(nd->flags); cudaFree(nd->x0); cudaFree(nd->zbuf); cudaFree(nd->head_partial); cudaFree(nd->head_count);
    nd->head_partial = nullptr; nd->head_count = nullptr;
    nd->planes = nullptr; nd->act[0] = nd->act[1] = nullptr; nd->flags = nullptr; nd->x0 = nullptr; nd->zbuf = nullptr;
    nd->cap = 0;
}

int ensure_workspace(NetDev* nd, int cap) {
    if (nd->cap >= cap) return LB2_OK;
    free_workspace(nd);
    nd->rows5 = round_up(cap * 441, 2 * lb2::kTileRows);  // whole CTA-pair items
    nd->rows3 = round_up(cap * 400, 2 * lb2::kTileRows);
    // the kernels index [planes][rows] buffers with 32-bit row offsets (activations: 2 * width / 8 planes; fused-head sums: 36)
    if ((uint64_t)std::max(2 * nd->width / 8, 9 * lb2::kColParts * lb2::kMaxSplit) * (uint64_t)nd->rows3 >= (1ull << 32))
        return fail(LB2_ERR_UNSUPPORTED, "batch of %d positions is too large for the activation workspace", cap);
    const size_t x0_bytes = (size_t)4 * nd->rows5 * 16;
    // fp16 planes + the extra planes of the split-operand modes (fp16 residuals, or e4m3 activations and residuals)
    const size_t act_bytes = (size_t)(2 * nd->width / 8) * nd->rows3 * 16;
    CU_TRY(cudaMalloc(&nd->planes, (size_t)cap * lb2::kPoints * sizeof(uint32_t)));
    CU_TRY(cudaMalloc(&nd->x0, x0_bytes));
    CU_TRY(cudaMemset(nd->x0, 0, x0_bytes));
    CU_TRY(cudaMalloc(&nd->zbuf, (size_t)9 * lb2::kColParts * lb2::kMaxSplit * nd->rows3 * sizeof(float)));
    if (nd->hidden) {
        CU_TRY(cudaMalloc(&nd->head_partial, lb2::heads_partial_floats(cap) * sizeof(float)));
        CU_TRY(cudaMalloc(&nd->head_count, lb2::heads_count_words(cap) * sizeof(uint32_t)));
        CU_TRY(cudaMemset(nd->head_count, 0, lb2::heads_count_words(cap) * sizeof(uint32_t)));
    }
    CU_TRY(cudaMalloc(&nd->act[0], act_bytes));
    CU_TRY(cudaMalloc(&nd->act[1], act_bytes));
    CU_TRY(cudaMemset(nd->act[0], 0, act_bytes));  // padding rows/columns must start (and stay) zero
    CU_TRY(cudaMemset(nd->act[1], 0, act_bytes));
    nd->flags_stride = nd->rows5 / lb2::kTileRows + 2;
    const size_t n_flags = (size_t)lb2::kMaxLayers * lb2::kMaxSplit * nd->flags_stride;
    CU_TRY(cudaMalloc(&nd->flags, n_flags * sizeof(uint32_t)));
    CU_TRY(cudaMemset(nd->flags, 0, n_flags * sizeof(uint32_t)));
    int rc;
    if ((rc = make_act_tmap(&nd->tm_x0, nd->x0, nd->rows5, 4, 48))) return rc;
    if ((rc = make_act_tmap(&nd->tm_act[0], nd->act[0], nd->rows3, 2 * nd->width / 8, 24))) return rc;
    if ((rc = make_act_tmap(&nd->tm_act[1], nd->act[1], nd->rows3, 2 * nd->width / 8, 24))) return rc;
    nd->cap = cap;
    return LB2_OK;
}

void drop_graphs(DeviceState* d) {
    for (auto& g : d->graphs) {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        if (g.graph) cudaGraphDestroy(g.graph);
    }
    d->graphs.clear();
    d->seen.clear();
}

// Grow the workspaces of the nets a call needs. Nothing may be using them meanwhile, and every cached graph holds
// pointers into the old ones. Caller holds d->mu.
int grow_workspaces(DeviceState* d, const bool need[2], int cap) {
    bool grow = false;
    for (int k = 0; k < 2; k++) grow = grow || (need[k] && d->net[k].cap < cap);
    if (!grow) return LB2_OK;
    CU_TRY(cudaDeviceSynchronize());
    drop_graphs(d);
    for (int k = 0; k < 2; k++) {
        int rc;
        if (need[k] && (rc = ensure_workspace(&d->net[k], cap))) return rc;
    }
    return LB2_OK;
}

// --------------------------------------------------------------------------------------------
// the batch pipeline on one device; all pointers are device pointers
// --------------------------------------------------------------------------------------------
struct Options {   // snapshot of the context's options for one call
    long trunk_mode, cta_pair, dynamic_items, use_graphs, resident_weights, policy_clusters, profile_trunk, max_batch, spin_wait, group_positions, small_batch, epoch;
    int prec[2];
};
Options snapshot(lb2_ctx* ctx) {
    std::lock_guard<std::mutex> lk(ctx->opt_mu);
    Options o;
    o.trunk_mode = ctx->trunk_mode; o.cta_pair = ctx->cta_pair; o.dynamic_items = ctx->dynamic_items; o.use_graphs = ctx->use_graphs;
    o.resident_weights = ctx->resident_weights; o.policy_clusters = ctx->policy_clusters; o.profile_trunk = ctx->profile_trunk;
    o.max_batch = ctx->max_batch; o.spin_wait = ctx->spin_wait; o.group_positions = ctx->group_positions; o.small_batch = ctx->small_batch; o.epoch = ctx->option_epoch.load();
    o.prec[0] = (int)ctx->precision[0]; o.prec[1] = (int)ctx->precision[1];
    return o;
}

struct JobPlan {
    std::vector<lb2::LayerJob> jobs;
    std::vector<int> round_of;  // launch round (layer depth) of each job, for per-layer mode
    std::vector<int> tiles;     // 256-row tiles of each job
    int total_items = 0;
    const __half* last_act[2] = {nullptr, nullptr};
    int tmap_base[2] = {0, 3};
    int net_begin[2] = {0, 0}, net_end[2] = {0, 0};   // net-major order: item index range of each net
    double net_cost[2] = {0, 0};                      // tensor-pipe cycles, for splitting the clusters between the nets
};

// Interleave the two nets layer by layer: P1 V1 P2a P2b V2 ... so that one launch round holds the
// independent jobs of equal depth (a layer wider than 128 channels contributes one job per column
// split; they are consecutive in the table).
// `net_major` (resident-weights mode): all policy jobs first, then all value jobs, every job a round of its own
JobPlan plan_jobs(DeviceState* d, const bool run[2], int n, const int limit_layers[2], bool pair, bool net_major, const int prec[2], bool small) {
    JobPlan pl;
    size_t depth = 0;
    for (int k = 0; k < 2; k++)
        if (run[k]) depth = std::max(depth, (size_t)std::min<int>(limit_layers[k], d->net[k].trunk.size()));
    int prev_job[2] = {-1, -1}, prev_split[2] = {1, 1};
    for (size_t outer = 0; outer < (net_major ? (size_t)2 : depth); outer++) {
        for (size_t inner = 0; inner < (net_major ? depth : (size_t)2); inner++) {
            const size_t l = net_major ? inner : outer;
            const int k = net_major ? (int)outer : (int)inner;
            if (net_major && inner == 0) pl.net_begin[k] = pl.net_end[k] = pl.total_items;
            NetDev& nd = d->net[k];
            if (!run[k] || l >= nd.trunk.size() || (int)l >= limit_layers[k]) continue;
            const TrunkLayerDev& t = nd.trunk[l];
            const int first_job = (int)pl.jobs.size();
            // small batches: every layer that has the packing runs as two column splits (CTA pairs only)
            const SplitPacks& S = (small && pair && t.has_small) ? t.small : t.nat;
            const int w = t.c_out / S.n_split;
            for (int sp = 0; sp < S.n_split; sp++) {
                lb2::LayerJob J;
                memset(&J, 0, sizeof J);
                const bool first = (l == 0);
                J.S = first ? 21 : 20;
                J.ksize = t.k;
                J.halo = first ? 48 : 24;
                const int mode = prec[k], n_real = t.c_in / 16, in_lo = t.c_in / 8;
                const bool consumed = l + 1 < nd.trunk.size() && (int)l + 1 < limit_layers[k];   // another trunk layer reads this one
                J.n_real_slabs = n_real;
                J.n_slabs = n_real * (mode == kPrecFp16 ? 1 : (mode == kPrecFull && !first ? 3 : 2));
                J.n_f16_slabs = (int16_t)(mode == kPrecLite && !first ? n_real : J.n_slabs);
                // input chunk planes of each term: hi | hi | lo (full), hi | e4m3 pairs (lite), hi | hi (first layer of either)
                J.term_base[0] = 0;
                J.term_base[1] = (int16_t)(mode == kPrecLite && !first ? in_lo : 0);
                J.term_base[2] = (int16_t)in_lo;
                J.lo_chunks = t.c_out / 8;   // the extra planes of a layer's output follow its c_out / 8 fp16 planes
                J.out_mode = !consumed || mode == kPrecFp16 ? lb2::kOutPlain : (mode == kPrecFull ? lb2::kOutLo16 : lb2::kOutFp8);
                J.acc_scale = mode == kPrecLite ? ldexpf(1.0f, -t.q) : 1.0f;
                J.n_out = w;
                const int n_tiles = (n * J.S * J.S + lb2::kTileRows - 1) / lb2::kTileRows;
                J.n_items = pair ? (n_tiles + 1) / 2 : n_tiles;
                J.item_base = pl.total_items;
                J.remap = first ? 1 : 0;
                // buffers: x0 -> act0 -> act1 -> act0 ...; a split writes its own chunk planes
                J.tmap = pl.tmap_base[k] + (first ? 0 : 1 + (int)((l - 1) & 1));
                J.out = nd.act[l & 1] + (size_t)(sp * w / 8) * nd.rows3 * 8;
                J.out_chunk_rows = nd.rows3;
                J.dep_job = prev_job[k];
                if (J.dep_job >= 0) {
                    J.dep_n_split = prev_split[k];
                    J.dep_remap = pl.jobs[J.dep_job].remap;
                    J.dep_n_items = pl.tiles[J.dep_job];
                }
                // the input tiles of this layer are dead once read: which planes the publisher may drop from L2
                if (l > 0 && S.n_split == 1 && prev_split[k] == 1) {
                    J.in_base = nd.act[(l - 1) & 1];
                    J.in_chunk_rows = nd.rows3;
                    J.in_planes = t.c_in / 8 * (mode == kPrecFp16 ? 1 : 2);
                }
                J.n_pos = n;
                J.net = k;
                J.layer = (int)l;
                J.wpk = S.w[mode][0][sp];
                J.wpk2 = S.w[mode][1][sp];
                J.bias = S.bias[sp];
                J.flags = nd.flags + ((size_t)l * lb2::kMaxSplit + sp) * nd.flags_stride;
                J.head_slot = net_major ? k : k * lb2::kMaxSplit + sp;
                if (l + 1 == nd.trunk.size() && limit_layers[k] > (int)nd.trunk.size()) {
                    // whole net: fold the final 3x3 conv to one channel into this layer's epilogue
                    J.head_taps = 9;
                    J.head_w = nd.head_wt[sp];
                    J.zbuf = nd.zbuf;
                    J.zparts = sp * lb2::kColParts;
                }
                pl.jobs.push_back(J);
                pl.round_of.push_back(net_major ? (int)l + 1000 * k : (int)l);
                {   // relative cost of an MMA by width, fitted to the best split measured (48 of 74 clusters on
                    // the policy net at batch 256): N = 128 runs at 64 cycles, narrower ones are bound by the A read
                    const double mma = J.n_out >= 128 ? 64.0 : (J.n_out > 64 ? 60.0 : 68.0);
                    pl.net_cost[k] += (double)J.n_items * J.n_slabs * t.k * t.k * 2 * mma;
                }
                pl.tiles.push_back(n_tiles);
                pl.total_items += J.n_items;
                if (net_major) pl.net_end[k] = pl.total_items;
                pl.last_act[k] = nd.act[l & 1];
            }
            prev_job[k] = first_job;
            prev_split[k] = S.n_split;
        }
    }
    return pl;
}

// The trunk launches of one evaluation (kernel parameters only; nothing is enqueued here).
int plan_trunk(const Options& o, DeviceState* d, const bool run[2], int n, const int limit_layers[2], EvalPlan* out) {
    const bool pair = o.cta_pair != 0 && d->sm_count >= 2;
    // resident-weights mode: CTA pairs, the single dataflow launch with dynamic claiming, and every layer's
    // half of the packed weights (all virtual slabs of its precision mode) must fit the resident area
    int modes = 0;   // split-operand modes present in the launch: selects the kernel instance
    for (int k = 0; k < 2; k++)
        if (run[k] && o.prec[k] != kPrecFp16) modes |= o.prec[k] == kPrecFull ? 1 : 2;
    if (modes == 3) return fail(LB2_ERR_UNSUPPORTED, "one net in lite and the other in full precision is not supported in one launch");
    for (int k = 0; k < 2; k++)   // (LB2_LITE_SEPARATE_ACC builds only: the e4m3 terms accumulate beside the fp16 sum in the same TMEM slot)
        if (lb2::kLiteSeparateAcc && run[k] && o.prec[k] == kPrecLite && d->net[k].width > lb2::kCorrCols)
            return fail(LB2_ERR_UNSUPPORTED, "lite precision needs layers of at most %d channels (the %s net has %d)", lb2::kCorrCols,
                        k == 0 ? "policy" : "value", d->net[k].width);
    // small batches run every layer as two column splits (SplitPacks small): twice the items, each half as long
    const bool small = pair && o.trunk_mode == 1 && n <= o.small_batch;
    const bool want_resident = o.resident_weights == 1 || (o.resident_weights == 2 && run[0] && run[1]);
    bool resident = pair && !small && want_resident && o.trunk_mode == 1 && o.dynamic_items != 0;
    int n_jobs_est = 0;
    for (int k = 0; k < 2 && resident; k++) {
        if (!run[k]) continue;
        for (size_t l = 0; l < d->net[k].trunk.size() && (int)l < limit_layers[k]; l++) {
            const TrunkLayerDev& t = d->net[k].trunk[l];
            const int terms = o.prec[k] == kPrecFp16 ? 1 : (o.prec[k] == kPrecFull && l > 0 ? 3 : 2);
            if (t.nat.n_split != 1 || (size_t)t.k * t.k * t.c_in * terms * (t.c_out / 2) * 2 > (size_t)lb2::kResWeightBytes) resident = false;
            n_jobs_est++;
        }
    }
    if (n_jobs_est > lb2::kResJobs) resident = false;
    JobPlan pl = plan_jobs(d, run, n, limit_layers, pair, resident, o.prec, small);
    if (pl.jobs.empty()) return fail(LB2_ERR_STATE, "no trunk layers to run");
    if ((int)pl.jobs.size() > lb2::kMaxLaunchJobs) return fail(LB2_ERR_UNSUPPORTED, "too many layers");
    TrunkLaunch L;
    lb2::TrunkParams& P = L.P;
    memset(&P, 0, sizeof P);
    L.pair = pair; L.resident = resident; L.modes = modes;
    for (int k = 0; k < 2; k++) {
        if (!d->net[k].cap) continue;
        P.tmaps[pl.tmap_base[k] + 0] = d->net[k].tm_x0;
        P.tmaps[pl.tmap_base[k] + 1] = d->net[k].tm_act[0];
        P.tmaps[pl.tmap_base[k] + 2] = d->net[k].tm_act[1];
    }
    for (int k = 0; k < 2; k++)  // unused slots still get prefetched: point them at a valid map
        if (!d->net[k].cap)
            for (int i = 0; i < 3; i++) P.tmaps[pl.tmap_base[k] + i] = d->net[1 - k].tm_x0;
    memcpy(P.jobs, pl.jobs.data(), pl.jobs.size() * sizeof(lb2::LayerJob));
    P.n_jobs = (int)pl.jobs.size();
    // Position groups (net-major order only): all layers of the first `group` positions, then of the next ... — the live
    // activations (a layer's input and output of ONE group per net) then fit in L2 instead of being written back to HBM
    // between layers. Groups are multiples of 128 positions: 128 x 400 rows = 100 whole CTA-pair items of an S = 20 layer.
    int n_groups = 1, group = 0;
    if (resident && o.group_positions > 0 && n > o.group_positions) {
        n_groups = (int)std::min<long>(4, (n + o.group_positions - 1) / o.group_positions);
        group = round_up((n + n_groups - 1) / n_groups, 128);
        n_groups = (n + group - 1) / group;
    }
    auto group_items = [&](const lb2::LayerJob& J) { return J.S == 21 ? (group * 441 + 511) / 512 : group * 400 / 512; };
    if (n_groups > 1)
        for (size_t i = 0; i < pl.jobs.size(); i++) {
            lb2::LayerJob& J = P.jobs[i];
            J.group_tiles = 2 * group_items(J);
            J.dep_group_tiles = J.dep_job >= 0 ? 2 * group_items(P.jobs[J.dep_job]) : 0;
        }
    // rounds: jobs of equal depth, their items interleaved in the launch-wide order; or (position groups) one group of one job
    P.n_rounds = 0;
    P.round_base[0] = 0;
    if (n_groups > 1) {
        for (int k = 0; k < 2; k++)
            for (int g = 0; g < n_groups; g++)
                for (size_t i = 0; i < pl.jobs.size(); i++) {
                    const lb2::LayerJob& J = P.jobs[i];
                    if (J.net != k) continue;
                    const int idx0 = g * group_items(J), cnt = std::min(group_items(J), J.n_items - idx0);
                    if (cnt <= 0) continue;
                    if (P.n_rounds >= lb2::kMaxRounds) return fail(LB2_ERR_UNSUPPORTED, "too many layers");
                    P.round_first[P.n_rounds] = (int16_t)i;
                    P.round_jobs[P.n_rounds] = 1;
                    P.round_idx0[P.n_rounds] = (int16_t)idx0;
                    P.round_base[P.n_rounds + 1] = P.round_base[P.n_rounds] + cnt;
                    P.n_rounds++;
                }
    } else {
        for (size_t i = 0; i < pl.jobs.size();) {
            size_t e = i;
            while (e < pl.jobs.size() && pl.round_of[e] == pl.round_of[i]) e++;
            if (P.n_rounds >= lb2::kMaxRounds || e - i > (size_t)lb2::kMaxRoundJobs) return fail(LB2_ERR_UNSUPPORTED, "too many layers");
            P.round_first[P.n_rounds] = (int16_t)i;
            P.round_jobs[P.n_rounds] = (int16_t)(e - i);
            int items = 0;
            for (size_t k = i; k < e; k++) items += pl.jobs[k].n_items;
            P.round_base[P.n_rounds + 1] = P.round_base[P.n_rounds] + items;
            P.n_rounds++;
            i = e;
        }
    }
    for (int r = P.n_rounds + 1; r <= lb2::kMaxRounds; r++) P.round_base[r] = 0x7fffffff;
    P.sched = d->sched;
    P.dynamic = o.dynamic_items ? 1 : 0;
    P.trace = d->trace;
#ifdef LB2_DEBUG_KNOBS
    if (const char* dbg = getenv("LB2_DEBUG_FLAGS")) P.debug_flags = atoi(dbg);
#endif
    out->trunk.clear();
    if (o.trunk_mode == 1) {
        P.item_begin = 0;
        P.item_end = pl.total_items;
        P.use_flags = 1;
        int grid = pair ? std::min(d->sm_count & ~1, 2 * pl.total_items) : std::min(d->sm_count, pl.total_items);
#ifdef LB2_DEBUG_KNOBS
        if (const char* g = getenv("LB2_GRID")) grid = std::max(2, std::min(grid, atoi(g) & ~1));   // experiment: fewer SMs (power-cap study)
#endif
        const int n_clusters = pair ? grid / 2 : grid;
        if (resident) {
            // each cluster draws items of its preferred net until that counter runs past the end, then helps the other
            for (int k = 0; k < 2; k++) {
                P.net_item_begin[k] = pl.net_begin[k];
                P.net_item_end[k] = pl.net_end[k];
            }
            const double total = pl.net_cost[0] + pl.net_cost[1];
            int pc = total > 0 ? (int)(n_clusters * pl.net_cost[0] / total + 0.5) : n_clusters;
            if (pl.net_cost[0] > 0 && pl.net_cost[1] > 0) pc = std::max(1, std::min(n_clusters - 1, pc));
            if (o.policy_clusters >= 0 && pl.net_cost[0] > 0 && pl.net_cost[1] > 0) pc = (int)std::min<long>(n_clusters - 1, std::max<long>(1, o.policy_clusters));
            P.policy_clusters = pc;
        }
        P.claim_base = 0;   // the expand kernel in front of the launch zeroes the claim counters
        L.grid = grid;
        L.cooperative = true;
        out->trunk.push_back(L);
    } else {
        // one launch per round; they share the claim counter, each advancing it by its items + one end marker per cluster
        P.use_flags = 0;
        uint32_t claimed = 0;
        for (int r = 0; r < P.n_rounds; r++) {
            P.item_begin = P.round_base[r];
            P.item_end = P.round_base[r + 1];
            const int items = P.item_end - P.item_begin;
            L.grid = pair ? std::min(d->sm_count & ~1, 2 * items) : std::min(d->sm_count, items);
            L.cooperative = false;
            L.resident = false;
            P.claim_base = claimed;
            claimed += (uint32_t)items + (uint32_t)(pair ? L.grid / 2 : L.grid);
            out->trunk.push_back(L);
        }
    }
    out->last_act[0] = pl.last_act[0];
    out->last_act[1] = pl.last_act[1];
    return LB2_OK;
}

// Every kernel argument of the evaluation `key` describes. Caller holds d->mu; the workspaces are large enough.
int build_plan(const Options& o, DeviceState* d, const EvalKey& key, EvalPlan* pl) {
    const bool run[2] = {key.limit[0] > 0, key.limit[1] > 0};
    const int n = key.n;
    float temp;
    memcpy(&temp, &key.temp_bits, sizeof temp);
    lb2::ExpandArgs& ea = pl->ea;
    memset(&ea, 0, sizeof ea);
    ea.rotation = static_cast<const uint8_t*>(key.in[2]);
    ea.ensemble = key.ensemble ? 1 : 0;
    ea.n = n;
    ea.sched = d->sched;
    for (int k = 0; k < 2; k++) {
        if (!run[k]) continue;
        NetDev& nd = d->net[k];
        if (nd.cap < n) return fail(LB2_ERR_STATE, "workspace too small");
        ea.planes[ea.n_nets] = static_cast<const uint32_t*>(key.in[k]);
        ea.x0[ea.n_nets] = nd.x0;
        ea.chunk_rows[ea.n_nets] = nd.rows5;
        ea.n_nets++;
        if (k == 1 && nd.ip1_wt && key.limit[1] > (int)nd.trunk.size()) {  // pull the value head's matrix into L2 while the trunk runs
            ea.pf = reinterpret_cast<const uint8_t*>(nd.ip1_wt);
            ea.pf_bytes = (size_t)lb2::kPoints * nd.hidden * sizeof(float);
        }
    }
    int rc = plan_trunk(o, d, run, n, key.limit, pl);
    if (rc) return rc;
    lb2::HeadArgs& ha = pl->ha;
    memset(&ha, 0, sizeof ha);
    ha.rotation = static_cast<const uint8_t*>(key.in[2]);
    ha.ensemble = key.ensemble ? 1 : 0;
    ha.temp = temp;
    pl->heads = false;
    if (run[0] && key.out[0] && key.limit[0] > (int)d->net[0].trunk.size()) {
        NetDev& nd = d->net[0];
        ha.p_zbuf = nd.zbuf; ha.p_chunk_rows = nd.rows3; ha.p_bias = nd.head_b; ha.probs = static_cast<float*>(key.out[0]); ha.n_policy = n;
        ha.p_parts = nd.trunk.back().nat.n_split * lb2::kColParts;
        pl->heads = true;
    }
    if (run[1] && key.out[1] && key.limit[1] > (int)d->net[1].trunk.size()) {
        NetDev& nd = d->net[1];
        ha.v_zbuf = nd.zbuf; ha.v_chunk_rows = nd.rows3; ha.v_bias = nd.head_b; ha.ip1_wt = nd.ip1_wt; ha.ip1_b = nd.ip1_b;
        ha.hidden = nd.hidden; ha.ip2_w = nd.ip2_w; ha.ip2_b = nd.ip2_b; ha.winrate = static_cast<float*>(key.out[1]); ha.n_value = n;
        ha.v_parts = nd.trunk.back().nat.n_split * lb2::kColParts;
        ha.v_partial = nd.head_partial; ha.v_count = nd.head_count;
        pl->heads = true;
    }
    ha.trace = d->trace;
    ha.trace_ctas = d->sm_count;
    memset(&pl->ma, 0, sizeof pl->ma);
    pl->mean = false;
    if (key.ensemble) {
        lb2::MeanArgs& ma = pl->ma;
        if (key.out[0]) { ma.probs8 = static_cast<float*>(key.out[0]); ma.probs = static_cast<float*>(key.mean[0]); ma.n_policy = n / 8; }
        if (key.out[1]) { ma.win8 = static_cast<float*>(key.out[1]); ma.win = static_cast<float*>(key.mean[1]); ma.n_value = n / 8; }
        pl->mean = true;
    }
    pl->kernels = 1 + (int)pl->trunk.size() + (pl->heads ? 1 : 0) + (pl->mean ? 1 : 0);
    return LB2_OK;
}

// (start, stop) event pair around the trunk of the evaluation being enqueued; from the pool if it has any left
int take_prof_events(DeviceState* d, cudaEvent_t* e0, cudaEvent_t* e1) {
    cudaEvent_t* ev[2] = {e0, e1};
    for (auto e : ev) {
        if (!d->prof_pool.empty()) { *e = d->prof_pool.back(); d->prof_pool.pop_back(); }
        else CU_TRY(cudaEventCreate(e));
    }
    d->prof_events.push_back(*e0);
    d->prof_events.push_back(*e1);
    return LB2_OK;
}

// Enqueue the kernels of a plan on `st`, one launch after the other.
// (`capturing`: the events become event-record NODES of the graph being captured, not capture-internal markers)
int launch_plan(const EvalPlan& pl, cudaStream_t st, cudaEvent_t ev0, cudaEvent_t ev1, bool capturing) {
    const unsigned ev_flags = capturing ? cudaEventRecordExternal : cudaEventRecordDefault;
    CU_TRY(lb2::launch_expand(pl.ea, st));
    if (ev0) CU_TRY(cudaEventRecordWithFlags(ev0, st, ev_flags));
    for (const TrunkLaunch& L : pl.trunk) CU_TRY(lb2::launch_trunk(L.P, L.grid, L.cooperative, L.pair, L.resident, L.modes, st));
    if (ev1) CU_TRY(cudaEventRecordWithFlags(ev1, st, ev_flags));
    if (pl.heads) CU_TRY(lb2::launch_heads(pl.ha, st));
    if (pl.mean) CU_TRY(lb2::launch_ensemble_mean(pl.ma, st));
    return LB2_OK;
}

// Capture the plan's launches into a graph and instantiate it. The kernel nodes whose arguments hold caller pointers
// (expand: inputs, heads / mean: outputs) and the two event-record nodes are remembered so that a later launch of the
// same shape with other buffers only patches those nodes.
int build_graph(DeviceState* d, const EvalKey& key, const EvalPlan& pl, cudaEvent_t ev0, cudaEvent_t ev1, GraphEntry* g) {
    cudaStream_t cap = nullptr;
    CU_TRY(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
    cudaError_t e = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) { cudaStreamDestroy(cap); return fail(LB2_ERR_CUDA, "cudaStreamBeginCapture: %s", cudaGetErrorString(e)); }
    int rc = launch_plan(pl, cap, ev0, ev1, true);
    cudaGraph_t graph = nullptr;
    e = cudaStreamEndCapture(cap, &graph);
    cudaStreamDestroy(cap);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess || !graph) return fail(LB2_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
    g->graph = graph;
    size_t n_nodes = 0;
    CU_TRY(cudaGraphGetNodes(graph, nullptr, &n_nodes));
    std::vector<cudaGraphNode_t> nodes(n_nodes);
    CU_TRY(cudaGraphGetNodes(graph, nodes.data(), &n_nodes));
    g->n_expand = g->n_heads = g->n_mean = g->n_ev0 = g->n_ev1 = nullptr;
    g->kernels = 0;
    for (auto nd : nodes) {
        cudaGraphNodeType ty;
        CU_TRY(cudaGraphNodeGetType(nd, &ty));
        if (ty == cudaGraphNodeTypeKernel) {
            cudaKernelNodeParams kp;
            CU_TRY(cudaGraphKernelNodeGetParams(nd, &kp));
            if (kp.func == lb2::kernel_address(0)) g->n_expand = nd;
            else if (kp.func == lb2::kernel_address(1)) g->n_heads = nd;
            else if (kp.func == lb2::kernel_address(2)) g->n_mean = nd;
            g->kernels++;
        } else if (ty == cudaGraphNodeTypeEventRecord) {
            cudaEvent_t ev;
            CU_TRY(cudaGraphEventRecordNodeGetEvent(nd, &ev));
            if (ev == ev0) g->n_ev0 = nd; else if (ev == ev1) g->n_ev1 = nd;
        }
    }
    if (!g->n_expand || (pl.heads && !g->n_heads) || (pl.mean && !g->n_mean) || (ev0 && (!g->n_ev0 || !g->n_ev1)))
        return fail(LB2_ERR_CUDA, "captured graph is missing a node");
    CU_TRY(cudaGraphInstantiate(&g->exec, graph, 0));
    g->key = key;
    g->plan = pl;
    return LB2_OK;
}

// Point the nodes of a cached graph at other caller buffers.
int patch_graph(GraphEntry* g, const EvalKey& key) {
    EvalPlan& pl = g->plan;
    int slot = 0;
    for (int k = 0; k < 2; k++)
        if (key.limit[k] > 0) pl.ea.planes[slot++] = static_cast<const uint32_t*>(key.in[k]);
    pl.ea.rotation = static_cast<const uint8_t*>(key.in[2]);
    pl.ha.rotation = static_cast<const uint8_t*>(key.in[2]);
    if (pl.ha.n_policy) pl.ha.probs = static_cast<float*>(key.out[0]);
    if (pl.ha.n_value) pl.ha.winrate = static_cast<float*>(key.out[1]);
    struct { cudaGraphNode_t node; void* arg; } upd[3] = {{g->n_expand, &pl.ea}, {g->n_heads, &pl.ha}, {g->n_mean, &pl.ma}};
    if (pl.mean) {
        if (key.out[0]) { pl.ma.probs8 = static_cast<float*>(key.out[0]); pl.ma.probs = static_cast<float*>(key.mean[0]); }
        if (key.out[1]) { pl.ma.win8 = static_cast<float*>(key.out[1]); pl.ma.win = static_cast<float*>(key.mean[1]); }
    }
    for (auto& u : upd) {
        if (!u.node) continue;
        cudaKernelNodeParams kp;
        CU_TRY(cudaGraphKernelNodeGetParams(u.node, &kp));
        void* args[1] = {u.arg};
        kp.kernelParams = args;
        kp.extra = nullptr;
        CU_TRY(cudaGraphExecKernelNodeSetParams(g->exec, u.node, &kp));
    }
    g->key = key;
    return LB2_OK;
}

constexpr int kMaxGraphs = 24;   // per device

std::vector<long> shape_id(const EvalKey& k, int tag) {
    return {tag, k.n, k.limit[0], k.limit[1], k.ensemble, k.profile, (long)k.temp_bits, k.in[0] != nullptr, k.in[1] != nullptr};
}

// One evaluation on one device: expand -> trunk -> heads (-> ensemble mean) on stream `st`, ordered behind whatever
// used the shared activation workspace before. `tag` separates the graph caches of the callers (I/O slot index, or -1
// for calls on device pointers, whose buffers change from call to call). Caller holds d->mu.
int eval_on_device(lb2_ctx* ctx, const Options& o, DeviceState* d, EvalKey key, int tag, cudaStream_t st) {
    const bool user_stream = st != d->stream;
    if (d->last_on_user_stream) {
        CU_TRY(cudaStreamWaitEvent(st, d->ev_last, 0));
    } else if (user_stream) {
        CU_TRY(cudaEventRecord(d->ev_last, d->stream));
        CU_TRY(cudaStreamWaitEvent(st, d->ev_last, 0));
    }
    if (d->graphs_epoch != o.epoch) {   // an option changed: the cached graphs were built for other kernels
        if (!d->graphs.empty()) CU_TRY(cudaDeviceSynchronize());
        drop_graphs(d);
        d->graphs_epoch = o.epoch;
    }
    key.profile = o.profile_trunk != 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int rc;
    if (key.profile && (rc = take_prof_events(d, &ev0, &ev1))) return rc;
    GraphEntry* g = nullptr;
    const bool graphs_ok = o.use_graphs && o.trunk_mode == 1 && !d->trace;
    if (graphs_ok)
        for (auto& e : d->graphs)
            if (e.tag == tag && e.key.same_shape(key)) { g = &e; break; }
    int kernels = 0;
    if (g) {
        if (!g->key.same_pointers(key) && (rc = patch_graph(g, key))) return rc;
        if (ev0) {
            CU_TRY(cudaGraphExecEventRecordNodeSetEvent(g->exec, g->n_ev0, ev0));
            CU_TRY(cudaGraphExecEventRecordNodeSetEvent(g->exec, g->n_ev1, ev1));
        }
        CU_TRY(cudaGraphLaunch(g->exec, st));
        g->last_use = ++d->use_counter;
        kernels = g->kernels;
        ctx->graph_launches++;
    } else {
        EvalPlan pl;
        if ((rc = build_plan(o, d, key, &pl))) return rc;
        bool launched = false;
        if (d->seen.size() > 4096) d->seen.clear();   // (a caller cycling through thousands of shapes: start counting afresh)
        if (graphs_ok && d->seen[shape_id(key, tag)]++ >= 1) {
            // second use of the shape: worth a graph. A full cache gives up its least recently used entry, unless even that
            // one was used a moment ago (many shapes in rotation: graphs would be built and thrown away all the time)
            bool room = (int)d->graphs.size() < kMaxGraphs;
            if (!room) {
                size_t victim = 0;
                for (size_t i = 1; i < d->graphs.size(); i++)
                    if (d->graphs[i].last_use < d->graphs[victim].last_use) victim = i;
                if (d->use_counter - d->graphs[victim].last_use > 4 * kMaxGraphs) {
                    CU_TRY(cudaStreamSynchronize(d->stream));   // it may still be executing
                    if (d->last_on_user_stream) CU_TRY(cudaEventSynchronize(d->ev_last));
                    cudaGraphExecDestroy(d->graphs[victim].exec);
                    cudaGraphDestroy(d->graphs[victim].graph);
                    d->graphs.erase(d->graphs.begin() + victim);
                    room = true;
                }
            }
            GraphEntry ne;
            if (room && (rc = build_graph(d, key, pl, ev0, ev1, &ne)) == LB2_OK) {
                ne.tag = tag;
                ne.last_use = ++d->use_counter;
                d->graphs.push_back(ne);
                CU_TRY(cudaGraphLaunch(d->graphs.back().exec, st));
                ctx->graph_launches++;
                launched = true;
            } else if (room) {
                cudaGetLastError();   // fall back to plain launches (the failure text stays in lb2_last_error)
            }
        }
        if (!launched && (rc = launch_plan(pl, st, ev0, ev1, false))) return rc;
        kernels = pl.kernels;
    }
    ctx->launches += kernels;
    if (user_stream) CU_TRY(cudaEventRecord(d->ev_last, st));
    d->last_on_user_stream = user_stream;
    return LB2_OK;
}

int check_ready(lb2_ctx* ctx, const bool need[2]) {
    if (!ctx) return fail(LB2_ERR_INVALID, "null context");
    for (int k = 0; k < 2; k++)
        if (need[k] && !(ctx->nets[k] && ctx->nets[k]->finalized))
            return fail(LB2_ERR_STATE, "%s net not finalized", k == 0 ? "policy" : "value");
    return LB2_OK;
}

// true when [p, p + bytes) is page-locked host memory the device can DMA from/to directly. Registered ranges
// (lb2_register_host_buffer, the submit queue's own buffers) are answered from a list; other pointers cost one
// cudaPointerGetAttributes the first time and are remembered by address. A stale answer is harmless: cudaMemcpyAsync
// accepts pageable memory too (it then stages and synchronises internally).
bool is_pinned(lb2_ctx* ctx, const void* p, size_t bytes) {
    if (!p) return false;
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    {
        std::lock_guard<std::mutex> lk(ctx->pin_mu);
        for (auto& r : ctx->pinned_ranges)
            if (a >= r.first && a + bytes <= r.second) return true;
        for (auto& c : ctx->pin_cache)
            if (c.first == a) return c.second;
    }
    cudaPointerAttributes at;
    bool pinned = false;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) cudaGetLastError();
    else pinned = at.type == cudaMemoryTypeHost;
    std::lock_guard<std::mutex> lk(ctx->pin_mu);
    if (ctx->pin_cache.size() >= 256) ctx->pin_cache.erase(ctx->pin_cache.begin(), ctx->pin_cache.begin() + 128);
    ctx->pin_cache.emplace_back(a, pinned);
    return pinned;
}

int check_rotations(const uint8_t* rot, int n) {
    for (int i = 0; i < n; i++)
        if (rot[i] > 7) return fail(LB2_ERR_INVALID, "rotation[%d] = %d out of range 0..7", i, (int)rot[i]);
    return LB2_OK;
}

void free_slot(IoSlot* sl) {
    for (int k = 0; k < 2; k++) { cudaFree(sl->d_planes[k]); cudaFreeHost(sl->h_planes[k]); }
    cudaFree(sl->d_rot); cudaFree(sl->d_probs); cudaFree(sl->d_win);
    cudaFreeHost(sl->h_rot); cudaFreeHost(sl->h_probs); cudaFreeHost(sl->h_win);
    if (sl->ev_in) cudaEventDestroy(sl->ev_in);
    if (sl->ev_done) cudaEventDestroy(sl->ev_done);
    if (sl->ev_out) cudaEventDestroy(sl->ev_out);
    if (sl->stream) cudaStreamDestroy(sl->stream);
    sl->~IoSlot();
    new (sl) IoSlot();
}

// Buffers of a slot for `cap` device positions. The owner of the slot calls this; growing drops the device's cached
// graphs of this slot's buffers (they hold the old pointers).
int ensure_slot(DeviceState* d, IoSlot* sl, int cap) {
    if (!sl->stream) {
        CU_TRY(cudaStreamCreateWithFlags(&sl->stream, cudaStreamNonBlocking));
        CU_TRY(cudaEventCreateWithFlags(&sl->ev_in, cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&sl->ev_done, cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&sl->ev_out, cudaEventDisableTiming | cudaEventBlockingSync));
    }
    if (sl->cap >= cap) return LB2_OK;
    {
        std::lock_guard<std::mutex> lk(d->mu);
        CU_TRY(cudaDeviceSynchronize());
        drop_graphs(d);
    }
    for (int k = 0; k < 2; k++) { cudaFree(sl->d_planes[k]); cudaFreeHost(sl->h_planes[k]); sl->d_planes[k] = nullptr; sl->h_planes[k] = nullptr; }
    cudaFree(sl->d_rot); cudaFree(sl->d_probs); cudaFree(sl->d_win);
    cudaFreeHost(sl->h_rot); cudaFreeHost(sl->h_probs); cudaFreeHost(sl->h_win);
    sl->d_rot = nullptr; sl->d_probs = sl->d_win = nullptr; sl->h_rot = nullptr; sl->h_probs = sl->h_win = nullptr;
    sl->cap = 0;
    const size_t pbytes = (size_t)cap * lb2::kPoints * sizeof(uint32_t);
    for (int k = 0; k < 2; k++) {
        CU_TRY(cudaMalloc(&sl->d_planes[k], pbytes));
        CU_TRY(cudaMallocHost(&sl->h_planes[k], pbytes));
    }
    CU_TRY(cudaMalloc(&sl->d_rot, cap));
    CU_TRY(cudaMalloc(&sl->d_probs, (size_t)cap * lb2::kPoints * sizeof(float)));
    CU_TRY(cudaMalloc(&sl->d_win, (size_t)cap * sizeof(float)));
    CU_TRY(cudaMallocHost(&sl->h_rot, cap));
    CU_TRY(cudaMallocHost(&sl->h_probs, (size_t)cap * lb2::kPoints * sizeof(float)));
    CU_TRY(cudaMallocHost(&sl->h_win, (size_t)cap * sizeof(float)));
    sl->cap = cap;
    return LB2_OK;
}

// A free (device, slot) pair, searched round robin from `prefer` (or the context's rotating start). Blocks until one
// is free when `wait`; otherwise returns false at once.
bool acquire_slot(lb2_ctx* ctx, int prefer, bool wait, int* dev_out, int* slot_out) {
    std::unique_lock<std::mutex> lk(ctx->slot_mu);
    const int ndev = (int)ctx->dev.size();
    const int start = prefer >= 0 ? prefer % ndev : (int)(ctx->next_dev++ % (unsigned)ndev);
    for (;;) {
        // least loaded first: a device with both slots free before one that is already working
        for (int want_free = kIoSlots; want_free >= 1; want_free--)
            for (int i = 0; i < ndev; i++) {
                DeviceState& d = *ctx->dev[(start + i) % ndev];
                int free_slots = 0, pick = -1;
                for (int s = 0; s < kIoSlots; s++)
                    if (!d.slots[s].busy) { free_slots++; if (pick < 0) pick = s; }
                if (free_slots >= want_free && pick >= 0) {
                    d.slots[pick].busy = true;
                    d.slots[pick].phase.store(1, std::memory_order_release);
                    *dev_out = (start + i) % ndev;
                    *slot_out = pick;
                    return true;
                }
            }
        if (!wait) return false;
        ctx->slot_cv.wait(lk);
    }
}
void release_slot(lb2_ctx* ctx, int dev, int slot) {
    {
        std::lock_guard<std::mutex> lk(ctx->slot_mu);
        ctx->dev[dev]->slots[slot].busy = false;
        ctx->dev[dev]->slots[slot].phase.store(0, std::memory_order_release);
    }
    ctx->slot_cv.notify_all();
}

// One chunk of a host-buffer call, from the moment its slot is acquired to the moment its results are in the caller's
// buffers.
struct Chunk {
    int dev = -1, slot = -1;
    int lo = 0, cnt = 0;   // input positions [lo, lo + cnt) of the call
};

struct HostCall {
    lb2_ctx* ctx;
    Options opt;
    const uint32_t* src[2];
    const uint8_t* rot;
    float temp;
    float *probs, *win;
    bool need[2], ensemble;
    bool pin_in[2], pin_rot, pin_probs, pin_win;
};

// Copies up, kernels, copies down for one chunk; returns as soon as everything is enqueued.
int enqueue_chunk(const HostCall& c, const Chunk& ch, int cap) {
    DeviceState* d = c.ctx->dev[ch.dev].get();
    IoSlot* sl = &d->slots[ch.slot];
    const int dev_per_pos = c.ensemble ? 8 : 1, slot_per_pos = c.ensemble ? 9 : 1;
    int rc;
    CU_TRY(cudaSetDevice(d->id));
    if ((rc = ensure_slot(d, sl, slot_per_pos * cap))) return rc;
    // staging copies of pageable caller buffers happen outside the device lock
    if (!c.ensemble && !c.pin_rot) memcpy(sl->h_rot, c.rot + ch.lo, ch.cnt);
    for (int k = 0; k < 2; k++)
        if (c.need[k] && !c.pin_in[k])
            memcpy(sl->h_planes[k], c.src[k] + (size_t)ch.lo * lb2::kPoints, (size_t)ch.cnt * lb2::kPoints * sizeof(uint32_t));
    const size_t pbytes = (size_t)ch.cnt * lb2::kPoints * sizeof(uint32_t);
    const int n_dev = dev_per_pos * ch.cnt;
    std::lock_guard<std::mutex> lk(d->mu);
    if ((rc = grow_workspaces(d, c.need, dev_per_pos * cap))) return rc;
    if (!c.ensemble) CU_TRY(cudaMemcpyAsync(sl->d_rot, c.pin_rot ? c.rot + ch.lo : sl->h_rot, ch.cnt, cudaMemcpyHostToDevice, sl->stream));
    for (int k = 0; k < 2; k++) {
        if (!c.need[k]) continue;
        const uint32_t* from = c.pin_in[k] ? c.src[k] + (size_t)ch.lo * lb2::kPoints : sl->h_planes[k];
        CU_TRY(cudaMemcpyAsync(sl->d_planes[k], from, pbytes, cudaMemcpyHostToDevice, sl->stream));
    }
    CU_TRY(cudaEventRecord(sl->ev_in, sl->stream));
    CU_TRY(cudaStreamWaitEvent(d->stream, sl->ev_in, 0));
    EvalKey key;
    key.n = n_dev;
    key.limit[0] = c.need[0] ? 1 << 20 : 0;
    key.limit[1] = c.need[1] ? 1 << 20 : 0;
    key.ensemble = c.ensemble;
    memcpy(&key.temp_bits, &c.temp, sizeof c.temp);
    key.in[0] = c.need[0] ? sl->d_planes[0] : nullptr;
    key.in[1] = c.need[1] ? sl->d_planes[1] : nullptr;
    key.in[2] = sl->d_rot;
    key.out[0] = c.need[0] ? sl->d_probs : nullptr;
    key.out[1] = c.need[1] ? sl->d_win : nullptr;
    // ensemble: the 8 * cnt per-symmetry results land in the first 8 * cnt entries of the slot's output buffers, their
    // means behind them
    const float* res_probs = sl->d_probs;
    const float* res_win = sl->d_win;
    if (c.ensemble) {
        if (c.need[0]) { key.mean[0] = sl->d_probs + (size_t)n_dev * lb2::kPoints; res_probs = static_cast<float*>(key.mean[0]); }
        if (c.need[1]) { key.mean[1] = sl->d_win + n_dev; res_win = static_cast<float*>(key.mean[1]); }
    }
    if ((rc = eval_on_device(c.ctx, c.opt, d, key, ch.slot, d->stream))) return rc;
    CU_TRY(cudaEventRecord(sl->ev_done, d->stream));
    sl->phase.store(2, std::memory_order_release);
    CU_TRY(cudaStreamWaitEvent(sl->stream, sl->ev_done, 0));
    if (c.need[0])
        CU_TRY(cudaMemcpyAsync(c.pin_probs ? c.probs + (size_t)ch.lo * lb2::kPoints : sl->h_probs, res_probs,
                               (size_t)ch.cnt * lb2::kPoints * sizeof(float), cudaMemcpyDeviceToHost, sl->stream));
    if (c.need[1])
        CU_TRY(cudaMemcpyAsync(c.pin_win ? c.win + ch.lo : sl->h_win, res_win, (size_t)ch.cnt * sizeof(float), cudaMemcpyDeviceToHost,
                               sl->stream));
    CU_TRY(cudaEventRecord(sl->ev_out, sl->stream));
    return LB2_OK;
}

// Wait for a chunk's results and hand its slot back.
int finish_chunk(const HostCall& c, const Chunk& ch) {
    DeviceState* d = c.ctx->dev[ch.dev].get();
    IoSlot* sl = &d->slots[ch.slot];
    int rc = LB2_OK;
    cudaError_t e;
    if (c.opt.spin_wait) {
        // poll, giving the core away between polls: lowest latency when cores are free, no starvation when they are not
        while ((e = cudaEventQuery(sl->ev_out)) == cudaErrorNotReady) std::this_thread::yield();
    } else {
        e = cudaEventSynchronize(sl->ev_out);   // blocking-sync event: the thread sleeps
    }
    if (e != cudaSuccess) rc = fail(LB2_ERR_CUDA, "evaluation failed on device %d: %s", d->id, cudaGetErrorString(e));
    if (rc == LB2_OK) {
        if (c.need[0] && !c.pin_probs) memcpy(c.probs + (size_t)ch.lo * lb2::kPoints, sl->h_probs, (size_t)ch.cnt * lb2::kPoints * sizeof(float));
        if (c.need[1] && !c.pin_win) memcpy(c.win + ch.lo, sl->h_win, (size_t)ch.cnt * sizeof(float));
    }
    release_slot(c.ctx, ch.dev, ch.slot);
    return rc;
}

// Checks a host-buffer call and fills in everything a chunk needs to know about it.
int prepare_host_call(lb2_ctx* ctx, const uint32_t* pol, const uint32_t* val, const uint8_t* rot, int n, float temp,
                      float* probs, float* win, bool ensemble, HostCall* c) {
    c->ctx = ctx;
    c->need[0] = probs != nullptr; c->need[1] = win != nullptr;
    int rc = check_ready(ctx, c->need);
    if (rc) return rc;
    if (n < 0) return fail(LB2_ERR_INVALID, "n < 0");
    if (n == 0) return LB2_OK;
    if ((!ensemble && !rot) || (c->need[0] && !pol) || (c->need[1] && !val)) return fail(LB2_ERR_INVALID, "null input pointer");
    if (c->need[0] && !(temp > 0.0f)) return fail(LB2_ERR_INVALID, "softmax temperature must be > 0");
    if (!ensemble && (rc = check_rotations(rot, n))) return rc;
    c->opt = snapshot(ctx);
    c->src[0] = pol; c->src[1] = val; c->rot = rot; c->temp = temp; c->probs = probs; c->win = win; c->ensemble = ensemble;
    // caller buffers that are already page-locked are used directly; pageable ones go through the slot's pinned staging
    const size_t pb = (size_t)n * lb2::kPoints * 4;
    c->pin_in[0] = c->need[0] && is_pinned(ctx, pol, pb);
    c->pin_in[1] = c->need[1] && is_pinned(ctx, val, pb);
    c->pin_rot = ensemble || is_pinned(ctx, rot, n);
    c->pin_probs = c->need[0] && is_pinned(ctx, probs, pb);
    c->pin_win = c->need[1] && is_pinned(ctx, win, (size_t)n * 4);
    return LB2_OK;
}

// Host-buffer evaluation. The call is cut into chunks of at most max_batch positions; every chunk is one device batch
// on whichever (device, slot) is free — whole batches per device. A caller blocks for a slot only while it holds none
// itself (two callers each holding one slot and waiting for a second would otherwise deadlock): with chunks of its own
// in flight it first collects the oldest.
int eval_host(lb2_ctx* ctx, const uint32_t* pol, const uint32_t* val, const uint8_t* rot, int n, float temp,
              float* probs, float* win, bool ensemble = false) {
    HostCall c;
    int rc = prepare_host_call(ctx, pol, val, rot, n, temp, probs, win, ensemble, &c);
    if (rc || n == 0) return rc;
    // an ensemble position occupies 8 device positions (+1 for its mean in the output buffers)
    const int chunk = (int)std::min<long>(ensemble ? std::max<long>(1, c.opt.max_batch / 8) : c.opt.max_batch, n);
    std::deque<Chunk> flying;
    int first_error = LB2_OK;
    std::string first_text;
    auto collect_oldest = [&]() {
        const int r = finish_chunk(c, flying.front());
        flying.pop_front();
        if (r && !first_error) { first_error = r; first_text = g_last_error; }
    };
    for (int lo = 0; lo < n && !first_error; lo += chunk) {
        Chunk ch;
        ch.lo = lo;
        ch.cnt = std::min(chunk, n - lo);
        while (!acquire_slot(ctx, -1, flying.empty(), &ch.dev, &ch.slot)) collect_oldest();
        if ((rc = enqueue_chunk(c, ch, chunk))) {
            first_error = rc; first_text = g_last_error;
            cudaSetDevice(ctx->dev[ch.dev]->id);
            cudaStreamSynchronize(ctx->dev[ch.dev]->slots[ch.slot].stream);
            release_slot(ctx, ch.dev, ch.slot);
            break;
        }
        flying.push_back(ch);
    }
    while (!flying.empty()) collect_oldest();
    if (first_error) g_last_error = first_text;
    return first_error;
}

// One batch of at most max_batch positions on a (device, slot) the caller already holds; the slot is released when
// the results are in `probs` / `win`.
int eval_host_on_slot(lb2_ctx* ctx, int dev, int slot, const uint32_t* pol, const uint32_t* val, const uint8_t* rot, int n, float temp,
                      float* probs, float* win) {
    HostCall c;
    int rc = prepare_host_call(ctx, pol, val, rot, n, temp, probs, win, false, &c);
    if (rc || n == 0 || n > c.opt.max_batch) {
        release_slot(ctx, dev, slot);
        return rc ? rc : (n == 0 ? LB2_OK : fail(LB2_ERR_INVALID, "batch larger than max_batch"));
    }
    Chunk ch;
    ch.dev = dev; ch.slot = slot; ch.lo = 0; ch.cnt = n;
    if ((rc = enqueue_chunk(c, ch, (int)std::max<long>(n, std::min<long>(c.opt.max_batch, ctx->q_cap))))) {
        cudaSetDevice(ctx->dev[dev]->id);
        cudaStreamSynchronize(ctx->dev[dev]->slots[slot].stream);
        release_slot(ctx, dev, slot);
        return rc;
    }
    return finish_chunk(c, ch);
}

// --------------------------------------------------------------------------------------------
// asynchronous submission: requests of many threads packed into device batches
// --------------------------------------------------------------------------------------------
// A submitter appends its planes straight into the pinned buffer of the open batch of its kind (one short critical
// section: reserve + 1.4 KB copy per position); kDispatchersPerDevice dispatcher threads per device take whatever has accumulated
// as soon as they are free (so the batch size follows the load), evaluate it as ONE device batch on their device
// (eval_host with the pinned buffers: no second staging copy) and hand the results out.
QueueBatch* new_queue_batch(lb2_ctx* ctx) {
    std::unique_ptr<QueueBatch> b(new QueueBatch);
    b->cap = ctx->q_cap;
    const size_t pb = (size_t)b->cap * lb2::kPoints * 4;
    if (cudaMallocHost(&b->planes, pb) != cudaSuccess || cudaMallocHost(&b->rot, b->cap) != cudaSuccess ||
        cudaMallocHost(&b->out, pb) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    {
        std::lock_guard<std::mutex> lk(ctx->pin_mu);
        for (auto r : {std::make_pair((void*)b->planes, pb), std::make_pair((void*)b->rot, (size_t)b->cap), std::make_pair((void*)b->out, pb)})
            ctx->pinned_ranges.emplace_back(reinterpret_cast<uintptr_t>(r.first), reinterpret_cast<uintptr_t>(r.first) + r.second);
    }
    ctx->q_all.push_back(std::move(b));
    return ctx->q_all.back().get();
}

// A dispatcher holds a free I/O slot of `dev` while the device's compute stream still runs the batch of the other slot:
// kernels of one device execute one after the other, and below ~90 positions a pass costs the same ~130 us whatever its size
// (12 chained layers), so carrying off the two or three requests that have arrived so far would only put a second
// latency-bound pass behind the first. Let them accumulate until the running pass is done (ev_done: its kernels, not its
// copy down) or a full batch is waiting; the copy up and the launch of the new batch then overlap the old one's copy down
// and callbacks. (One request outstanding per thread, one B200: 16 threads 46 k -> 58 k requests/s, 64 threads 163 k -> 201 k,
// 128 threads 278 k -> 371 k: profiles/r2_queue_n1.json against r2_queue_n1_nolinger.json.)
void linger_for_device(lb2_ctx* ctx, int dev, int slot) {
    static_assert(kIoSlots == 2, "the other slot");
    IoSlot& mine = ctx->dev[dev]->slots[slot];
    IoSlot& other = ctx->dev[dev]->slots[slot ^ 1];
    // phase 3 = "waiting here": two dispatchers that took the two slots of an idle device at the same moment must not
    // wait for each other — whoever sees the other one lingering (3) or free (0) goes ahead
    mine.phase.store(3, std::memory_order_seq_cst);
    for (;;) {
        const int ph = other.phase.load(std::memory_order_seq_cst);
        if (ph == 0 || ph == 3) break;
        if (ph == 2 && cudaEventQuery(other.ev_done) != cudaErrorNotReady) break;
        if (ctx->q_waiting.load(std::memory_order_relaxed) >= ctx->q_cap) break;
        std::this_thread::yield();
    }
    mine.phase.store(1, std::memory_order_seq_cst);
}

void worker_loop(lb2_ctx* ctx, int dev_index) {
    cudaSetDevice(ctx->dev[dev_index]->id);
    auto have_work = [&] { return !ctx->q_ready.empty() || (ctx->q_open[0] && ctx->q_open[0]->n) || (ctx->q_open[1] && ctx->q_open[1]->n); };
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(ctx->q_mu);
            ctx->q_idle_workers++;
            ctx->q_cv.wait(lk, [&] { return !ctx->worker_run || have_work(); });
            ctx->q_idle_workers--;
            if (!have_work()) return;   // shutting down
        }
        // The slot FIRST, the batch second: while every I/O slot is busy, requests keep accumulating in the open batch
        // instead of being carried off in small batches that then queue for the device. (A small batch costs the device as
        // much time as a medium one: below ~90 positions a pass is bound by the latency of its 12 chained layers.)
        int dev = -1, slot = -1;
        acquire_slot(ctx, dev_index, true, &dev, &slot);
        bool linger;
        { std::lock_guard<std::mutex> ol(ctx->opt_mu); linger = ctx->queue_linger != 0; }
        cudaSetDevice(ctx->dev[dev]->id);   // (the slot may be another device's than this dispatcher's usual one)
        if (linger) linger_for_device(ctx, dev, slot);
        QueueBatch* b = nullptr;
        {
            std::lock_guard<std::mutex> lk(ctx->q_mu);
            if (!ctx->q_ready.empty()) {
                b = ctx->q_ready.front();
                ctx->q_ready.pop_front();
            } else {
                // nothing sealed: take the fuller of the open batches as it is
                int k = -1;
                for (int i = 0; i < 2; i++)
                    if (ctx->q_open[i] && ctx->q_open[i]->n && (k < 0 || ctx->q_open[i]->n > ctx->q_open[k]->n)) k = i;
                if (k >= 0) { b = ctx->q_open[k]; ctx->q_open[k] = nullptr; }
            }
            if (b) { ctx->workers_busy++; ctx->q_waiting.fetch_sub(b->n, std::memory_order_relaxed); }
        }
        if (!b) {   // another dispatcher took it meanwhile
            release_slot(ctx, dev, slot);
            continue;
        }
        cudaSetDevice(ctx->dev[dev]->id);
        ctx->stat_positions += b->n; ctx->stat_batches++; ctx->stat_requests += (long)b->req.size();
        const int rc = b->kind == LB2_POLICY ? eval_host_on_slot(ctx, dev, slot, b->planes, nullptr, b->rot, b->n, b->temp, b->out, nullptr)
                                             : eval_host_on_slot(ctx, dev, slot, nullptr, b->planes, b->rot, b->n, 1.0f, nullptr, b->out);
        const size_t per = b->kind == LB2_POLICY ? lb2::kPoints : 1;
        size_t o = 0;
        for (auto& r : b->req) {
            if (rc == LB2_OK) memcpy(r.out, b->out + o * per, (size_t)r.n * per * sizeof(float));
            o += r.n;
        }
        if (rc != LB2_OK) {
            std::lock_guard<std::mutex> lk(ctx->q_mu);
            ctx->q_error = g_last_error;
        }
        for (auto& r : b->req)
            if (r.cb) r.cb(r.user, rc);
        {
            std::lock_guard<std::mutex> lk(ctx->q_mu);
            b->n = 0;
            b->req.clear();
            ctx->q_free_list.push_back(b);
            ctx->workers_busy--;
            if (!have_work() && ctx->workers_busy == 0) ctx->q_idle.notify_all();
        }
        ctx->q_free.notify_one();
    }
}

int submit(lb2_ctx* ctx, int kind, const uint32_t* planes, const uint8_t* rot, int n, float temp, float* out,
           lb2_callback cb, void* user) {
    bool need[2] = {kind == LB2_POLICY, kind == LB2_VALUE};
    int rc = check_ready(ctx, need);
    if (rc) return rc;
    if (n <= 0 || !planes || !rot || !out) return fail(LB2_ERR_INVALID, "bad submit arguments");
    if ((rc = check_rotations(rot, n))) return rc;
    std::unique_lock<std::mutex> lk(ctx->q_mu);
    if (!ctx->worker_run) {
        ctx->worker_run = true;
        { std::lock_guard<std::mutex> ol(ctx->opt_mu); ctx->q_cap = (int)std::max<long>(ctx->max_batch, 1); }
        // more dispatchers than I/O slots: while kIoSlots of them have a batch on the device, the others hand out the
        // results of finished batches (a copy and a callback per request) and pick up the next
        for (size_t di = 0; di < ctx->dev.size(); di++)
            for (int i = 0; i < kDispatchersPerDevice; i++) ctx->workers.emplace_back(worker_loop, ctx, (int)di);
    }
    if (n > ctx->q_cap) {
        // larger than a batch buffer: evaluate it on its own (blocking), then report through the callback as usual
        lk.unlock();
        rc = kind == LB2_POLICY ? eval_host(ctx, planes, nullptr, rot, n, temp, out, nullptr)
                                : eval_host(ctx, nullptr, planes, rot, n, 1.0f, nullptr, out);
        if (cb) cb(user, rc);
        return LB2_OK;
    }
    for (;;) {
        QueueBatch*& open = ctx->q_open[kind];
        if (open && (open->n + n > open->cap || (open->n && open->temp != temp))) {   // full, or another temperature: seal it
            ctx->q_ready.push_back(open);
            open = nullptr;
        }
        if (!open) {
            if (!ctx->q_free_list.empty()) {
                open = ctx->q_free_list.back();
                ctx->q_free_list.pop_back();
            } else if (ctx->q_all.size() < 2 * kDispatchersPerDevice * ctx->dev.size() + 4) {
                if (!(open = new_queue_batch(ctx))) return fail(LB2_ERR_NOMEM, "pinned batch buffer");
            } else {
                ctx->q_free.wait(lk);   // every buffer is filling or in flight: back-pressure on the submitters
                continue;
            }
            open->kind = kind;
            open->temp = temp;
            open->n = 0;
        }
        memcpy(open->planes + (size_t)open->n * lb2::kPoints, planes, (size_t)n * lb2::kPoints * sizeof(uint32_t));
        memcpy(open->rot + open->n, rot, n);
        open->n += n;
        ctx->q_waiting.fetch_add(n, std::memory_order_relaxed);
        open->req.push_back(QueueRequest{out, n, cb, user});
        break;
    }
    const bool wake = ctx->q_idle_workers > 0;   // dispatchers that are busy (or waiting for a slot) look at the queue again by themselves
    lk.unlock();
    if (wake) ctx->q_cv.notify_one();
    return LB2_OK;
}

}  // namespace

// ============================================================================================
// C ABI
// ============================================================================================
extern "C" {

const char* lb2_last_error(void) { return g_last_error.c_str(); }

int lb2_init(const int* device_ids, int n_devices, lb2_ctx** ctx_out) {
    if (!ctx_out) return fail(LB2_ERR_INVALID, "ctx_out is null");
    *ctx_out = nullptr;
    int count = 0;
    CU_TRY(cudaGetDeviceCount(&count));
    if (count <= 0) return fail(LB2_ERR_CUDA, "no CUDA device");
    std::vector<int> ids;
    if (!device_ids || n_devices <= 0) ids.push_back(0);
    else ids.assign(device_ids, device_ids + n_devices);
    std::unique_ptr<lb2_ctx> ctx(new lb2_ctx);
    for (int id : ids) {
        if (id < 0 || id >= count) return fail(LB2_ERR_INVALID, "device id %d out of range (have %d)", id, count);
        cudaDeviceProp prop;
        CU_TRY(cudaGetDeviceProperties(&prop, id));
        if (prop.major != 10)
            return fail(LB2_ERR_UNSUPPORTED, "device %d (%s, sm_%d%d) is not a Blackwell B200-class GPU; no fallback path",
                        id, prop.name, prop.major, prop.minor);
        CU_TRY(cudaSetDevice(id));
        std::unique_ptr<DeviceState> d(new DeviceState);
        d->id = id;
        d->sm_count = prop.multiProcessorCount;
        CU_TRY(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
        CU_TRY(cudaMalloc(&d->sched, lb2::kSchedWords * sizeof(uint32_t)));
        CU_TRY(cudaMemset(d->sched, 0, lb2::kSchedWords * sizeof(uint32_t)));
        CU_TRY(cudaEventCreateWithFlags(&d->ev_last, cudaEventDisableTiming));
        CU_TRY(lb2::trunk_kernel_setup());
        if (ctx->backend.empty()) ctx->backend = std::string("B200 tcgen05: ") + prop.name;
        ctx->dev.push_back(std::move(d));
    }
    int rc = load_driver_entry();
    if (rc) return rc;
    if (ids.size() > 1) ctx->backend += " x" + std::to_string(ids.size());
    *ctx_out = ctx.release();
    return LB2_OK;
}

void lb2_destroy(lb2_ctx* ctx) {
    if (!ctx) return;
    lb2_drain(ctx);
    {
        std::lock_guard<std::mutex> lk(ctx->q_mu);
        ctx->worker_run = false;
    }
    ctx->q_cv.notify_all();
    for (auto& w : ctx->workers) if (w.joinable()) w.join();
    for (auto& b : ctx->q_all) { cudaFreeHost(b->planes); cudaFreeHost(b->rot); cudaFreeHost(b->out); }
    for (void* p : ctx->registered) cudaHostUnregister(p);
    for (auto& dp : ctx->dev) {
        DeviceState& d = *dp;
        cudaSetDevice(d.id);
        cudaDeviceSynchronize();
        drop_graphs(&d);
        for (int k = 0; k < 2; k++) {
            NetDev& nd = d.net[k];
            for (auto& t : nd.trunk)
                for (SplitPacks* S : {&t.nat, &t.small})
                    for (int sp = 0; sp < lb2::kMaxSplit; sp++) {
                        for (int mode = 0; mode < 3; mode++) { cudaFree(S->w[mode][0][sp]); cudaFree(S->w[mode][1][sp]); }
                        cudaFree(S->bias[sp]);
                    }
            for (int sp = 0; sp < lb2::kMaxSplit; sp++) cudaFree(nd.head_wt[sp]);
            cudaFree(nd.head_b); cudaFree(nd.ip1_wt); cudaFree(nd.ip1_b);
            cudaFree(nd.ip2_w); cudaFree(nd.ip2_b);
            free_workspace(&nd);
        }
        for (auto& sl : d.slots) free_slot(&sl);
        for (auto e : d.prof_events) cudaEventDestroy(e);
        for (auto e : d.prof_pool) cudaEventDestroy(e);
        if (d.ev_last) cudaEventDestroy(d.ev_last);
        cudaFree(d.rot); cudaFree(d.sched); cudaFree(d.trace);
        cudaStreamDestroy(d.stream);
    }
    delete ctx;
}

int lb2_net_create(lb2_ctx* ctx, int kind, lb2_net** net_out) {
    if (!ctx || !net_out) return fail(LB2_ERR_INVALID, "null argument");
    if (kind != LB2_POLICY && kind != LB2_VALUE) return fail(LB2_ERR_INVALID, "unknown net kind %d", kind);
    if (ctx->nets[kind]) return fail(LB2_ERR_STATE, "net of kind %d already exists", kind);
    ctx->nets[kind].reset(new lb2_net);
    ctx->nets[kind]->ctx = ctx;
    ctx->nets[kind]->kind = kind;
    *net_out = ctx->nets[kind].get();
    return LB2_OK;
}

int lb2_net_push_conv(lb2_net* net, int k, int c_in, int c_out, const float* w, const float* bias) {
    if (!net || !w || !bias) return fail(LB2_ERR_INVALID, "null argument");
    if (net->finalized) return fail(LB2_ERR_STATE, "net already finalized");
    if (!net->ips.empty()) return fail(LB2_ERR_STATE, "convolutions must precede inner products");
    if ((k != 3 && k != 5) || c_in <= 0 || c_out <= 0) return fail(LB2_ERR_INVALID, "bad conv geometry k=%d %d->%d", k, c_in, c_out);
    if (!net->convs.empty() && net->convs.back().c_out != c_in)
        return fail(LB2_ERR_INVALID, "conv input channels %d do not match previous output %d", c_in, net->convs.back().c_out);
    HostConv c;
    c.k = k; c.c_in = c_in; c.c_out = c_out;
    c.w.assign(w, w + (size_t)k * k * c_in * c_out);
    c.b.assign(bias, bias + c_out);
    net->convs.push_back(std::move(c));
    return LB2_OK;
}

int lb2_net_push_ip(lb2_net* net, int n_in, int n_out, const float* w, const float* bias) {
    if (!net || !w || !bias) return fail(LB2_ERR_INVALID, "null argument");
    if (net->finalized) return fail(LB2_ERR_STATE, "net already finalized");
    if (n_in <= 0 || n_out <= 0) return fail(LB2_ERR_INVALID, "bad inner product geometry");
    HostIp p;
    p.n_in = n_in; p.n_out = n_out;
    p.w.assign(w, w + (size_t)n_in * n_out);
    p.b.assign(bias, bias + n_out);
    net->ips.push_back(std::move(p));
    return LB2_OK;
}

int lb2_net_finalize(lb2_net* net) {
    if (!net) return fail(LB2_ERR_INVALID, "null net");
    if (net->finalized) return fail(LB2_ERR_STATE, "net already finalized");
    const auto& cv = net->convs;
    if (cv.size() < 3) return fail(LB2_ERR_UNSUPPORTED, "need at least 3 conv layers");
    if (cv[0].k != 5 || cv[0].c_in != LB2_INPUT_PLANES) return fail(LB2_ERR_UNSUPPORTED, "first layer must be 5x5 from 32 planes");
    for (size_t l = 0; l + 1 < cv.size(); l++) {
        if (l >= (size_t)lb2::kMaxLayers) return fail(LB2_ERR_UNSUPPORTED, "too many layers");
        if (l > 0 && cv[l].k != 3) return fail(LB2_ERR_UNSUPPORTED, "layer %zu: only 3x3 after the first layer", l + 1);
        // widths up to 128 run as one job (MMA N = c_out), up to 256 as two column splits of c_out / 2
        const bool ok_out = cv[l].c_out <= 128 ? cv[l].c_out % 32 == 0 : (cv[l].c_out <= 256 && cv[l].c_out % 64 == 0);
        if (!ok_out || cv[l].c_in % 16 || cv[l].c_in > 256)
            return fail(LB2_ERR_UNSUPPORTED, "layer %zu: channels %d->%d not supported", l + 1, cv[l].c_in, cv[l].c_out);
    }
    if (cv.back().k != 3 || cv.back().c_out != 1) return fail(LB2_ERR_UNSUPPORTED, "last conv must be 3x3 to 1 channel");
    if (net->kind == LB2_POLICY) {
        if (!net->ips.empty()) return fail(LB2_ERR_UNSUPPORTED, "policy net takes no inner products");
    } else {
        if (net->ips.size() != 2 || net->ips[0].n_in != LB2_BOARD_POINTS || net->ips[0].n_out > 256 ||
            net->ips[1].n_in != net->ips[0].n_out || net->ips[1].n_out != 1)
            return fail(LB2_ERR_UNSUPPORTED, "value net needs inner products 361->H (H<=256) and H->1");
        // the value head streams the 361 x H matrix in bulk copies of whole row tiles: sizes must be multiples of 16 bytes
        if (net->ips[0].n_out % 4) return fail(LB2_ERR_UNSUPPORTED, "value net hidden size %d is not a multiple of 4", net->ips[0].n_out);
    }
    for (auto& d : net->ctx->dev) {
        CU_TRY(cudaSetDevice(d->id));
        int rc = upload_net(net, &d->net[net->kind]);
        if (rc) return rc;
    }
    net->finalized = true;
    return LB2_OK;
}

int lb2_eval_policy(lb2_ctx* ctx, const uint32_t* planes, const uint8_t* rotation, int n, float temp, float* probs) {
    if (!probs && n > 0) return fail(LB2_ERR_INVALID, "null output pointer");
    return eval_host(ctx, planes, nullptr, rotation, n, temp, probs, nullptr);
}

int lb2_eval_value(lb2_ctx* ctx, const uint32_t* planes, const uint8_t* rotation, int n, float* winrate) {
    if (!winrate && n > 0) return fail(LB2_ERR_INVALID, "null output pointer");
    return eval_host(ctx, nullptr, planes, rotation, n, 1.0f, nullptr, winrate);
}

int lb2_eval_both(lb2_ctx* ctx, const uint32_t* pol, const uint32_t* val, const uint8_t* rotation, int n, float temp,
                  float* probs, float* winrate) {
    if ((!probs || !winrate) && n > 0) return fail(LB2_ERR_INVALID, "null output pointer");
    return eval_host(ctx, pol, val, rotation, n, temp, probs, winrate);
}

int lb2_eval_positions(lb2_ctx* ctx, const lb2_position* pos, const uint8_t* rotation, int n, float temp, float* probs,
                       float* winrate) {
    if (!probs && !winrate && n > 0) return fail(LB2_ERR_INVALID, "null output pointers");
    if (n < 0) return fail(LB2_ERR_INVALID, "n < 0");
    if (n == 0) return LB2_OK;
    if (!pos) return fail(LB2_ERR_INVALID, "null input pointer");
    std::vector<uint32_t> pol(probs ? (size_t)n * lb2::kPoints : 0), val(winrate ? (size_t)n * lb2::kPoints : 0);
    std::vector<uint8_t> rot(n, 0);
    if (rotation) rot.assign(rotation, rotation + n);
    // feature planes on the host's cores: ~45 us per middle-game position each
    const int n_threads = std::max(1, std::min<int>({(int)std::thread::hardware_concurrency(), 64, (n + 7) / 8}));
    std::atomic<int> next{0}, bad{0};
    auto work = [&]() {
        for (int i; (i = next.fetch_add(1)) < n;) {
            const lb2_position& p = pos[i];
            if (lb2_planes_from_position(p.stones, p.white_to_move, p.ko_point, p.last_move, p.prev_move, p.komi,
                                         probs ? &pol[(size_t)i * lb2::kPoints] : nullptr, winrate ? &val[(size_t)i * lb2::kPoints] : nullptr))
                bad++;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; t++) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    if (bad) return fail(LB2_ERR_INVALID, "%d invalid position(s) (stone values must be 0..2, point indices < 361)", bad.load());
    return eval_host(ctx, probs ? pol.data() : nullptr, winrate ? val.data() : nullptr, rot.data(), n, probs ? temp : 1.0f, probs, winrate);
}

int lb2_eval_ensemble(lb2_ctx* ctx, const uint32_t* pol, const uint32_t* val, int n, float temp, float* probs, float* winrate) {
    if (!probs && !winrate && n > 0) return fail(LB2_ERR_INVALID, "null output pointers");
    return eval_host(ctx, probs ? pol : nullptr, winrate ? val : nullptr, nullptr, n, probs ? temp : 1.0f, probs, winrate, true);
}

int lb2_eval_both_device(lb2_ctx* ctx, int dev_index, const uint32_t* d_pol, const uint32_t* d_val,
                         const uint8_t* d_rot, int n, float temp, float* d_probs, float* d_win, void* stream) {
    bool need[2] = {d_probs != nullptr, d_win != nullptr};
    int rc = check_ready(ctx, need);
    if (rc) return rc;
    if (dev_index < 0 || dev_index >= (int)ctx->dev.size()) return fail(LB2_ERR_INVALID, "bad device index");
    if (n < 0) return fail(LB2_ERR_INVALID, "n < 0");
    if (n == 0) return LB2_OK;
    if (!d_rot || (need[0] && !d_pol) || (need[1] && !d_val)) return fail(LB2_ERR_INVALID, "null input pointer");
    if (need[0] && !(temp > 0.0f)) return fail(LB2_ERR_INVALID, "softmax temperature must be > 0");
    const Options o = snapshot(ctx);
    DeviceState* d = ctx->dev[dev_index].get();
    std::lock_guard<std::mutex> lk(d->mu);
    CU_TRY(cudaSetDevice(d->id));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : d->stream;
    const int chunk = (int)o.max_batch;
    if ((rc = grow_workspaces(d, need, std::min(n, chunk)))) return rc;
    for (int lo = 0; lo < n; lo += chunk) {
        EvalKey key;
        key.n = std::min(chunk, n - lo);
        key.limit[0] = need[0] ? 1 << 20 : 0;
        key.limit[1] = need[1] ? 1 << 20 : 0;
        memcpy(&key.temp_bits, &temp, sizeof temp);
        key.in[0] = need[0] ? d_pol + (size_t)lo * lb2::kPoints : nullptr;
        key.in[1] = need[1] ? d_val + (size_t)lo * lb2::kPoints : nullptr;
        key.in[2] = d_rot + lo;
        key.out[0] = need[0] ? d_probs + (size_t)lo * lb2::kPoints : nullptr;
        key.out[1] = need[1] ? d_win + lo : nullptr;
        if ((rc = eval_on_device(ctx, o, d, key, -1, st))) return rc;
    }
    return LB2_OK;
}

int lb2_submit_policy(lb2_ctx* ctx, const uint32_t* planes, const uint8_t* rotation, int n, float temp, float* probs,
                      lb2_callback cb, void* user) {
    if (!(temp > 0.0f)) return fail(LB2_ERR_INVALID, "softmax temperature must be > 0");
    return submit(ctx, LB2_POLICY, planes, rotation, n, temp, probs, cb, user);
}

int lb2_submit_value(lb2_ctx* ctx, const uint32_t* planes, const uint8_t* rotation, int n, float* winrate,
                     lb2_callback cb, void* user) {
    return submit(ctx, LB2_VALUE, planes, rotation, n, 1.0f, winrate, cb, user);
}

int lb2_drain(lb2_ctx* ctx) {
    if (!ctx) return fail(LB2_ERR_INVALID, "null context");
    std::unique_lock<std::mutex> lk(ctx->q_mu);
    ctx->q_idle.wait(lk, [&] {
        return ctx->q_ready.empty() && !(ctx->q_open[0] && ctx->q_open[0]->n) && !(ctx->q_open[1] && ctx->q_open[1]->n) && ctx->workers_busy == 0;
    });
    return LB2_OK;
}

int lb2_queue_error(lb2_ctx* ctx, char* buf, int len) {
    if (!ctx || !buf || len <= 0) return fail(LB2_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> lk(ctx->q_mu);
    snprintf(buf, (size_t)len, "%s", ctx->q_error.c_str());
    return LB2_OK;
}

int lb2_register_host_buffer(lb2_ctx* ctx, void* ptr, size_t bytes) {
    if (!ctx || !ptr || !bytes) return fail(LB2_ERR_INVALID, "bad arguments");
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) cudaGetLastError();   // somebody else page-locked it: just remember the range
    else if (e != cudaSuccess) { cudaGetLastError(); return fail(LB2_ERR_CUDA, "cudaHostRegister failed: %s", cudaGetErrorString(e)); }
    std::lock_guard<std::mutex> lk(ctx->pin_mu);
    if (e == cudaSuccess) ctx->registered.push_back(ptr);
    const uintptr_t a = reinterpret_cast<uintptr_t>(ptr);
    ctx->pinned_ranges.emplace_back(a, a + bytes);
    ctx->pin_cache.clear();
    return LB2_OK;
}

int lb2_unregister_host_buffer(lb2_ctx* ctx, void* ptr) {
    if (!ctx || !ptr) return fail(LB2_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> lk(ctx->pin_mu);
    const uintptr_t a = reinterpret_cast<uintptr_t>(ptr);
    bool found = false;
    for (size_t i = 0; i < ctx->pinned_ranges.size(); i++)
        if (ctx->pinned_ranges[i].first == a) { ctx->pinned_ranges.erase(ctx->pinned_ranges.begin() + i); found = true; break; }
    if (!found) return fail(LB2_ERR_INVALID, "buffer was not registered");
    for (size_t i = 0; i < ctx->registered.size(); i++)
        if (ctx->registered[i] == ptr) { cudaHostUnregister(ptr); ctx->registered.erase(ctx->registered.begin() + i); break; }
    ctx->pin_cache.clear();
    return LB2_OK;
}

const char* lb2_backend_name(lb2_ctx* ctx) { return ctx ? ctx->backend.c_str() : ""; }
int lb2_device_count(lb2_ctx* ctx) { return ctx ? (int)ctx->dev.size() : 0; }

int lb2_set_option(lb2_ctx* ctx, const char* name, long value) {
    if (!ctx || !name) return fail(LB2_ERR_INVALID, "null argument");
    if (!strcmp(name, "trace")) {
        // debug: per-item timeline of the trunk kernel on device 0 (read back with lb2_debug_read_trace)
        DeviceState& d = *ctx->dev[0];
        std::lock_guard<std::mutex> lk(d.mu);
        cudaSetDevice(d.id);
        cudaDeviceSynchronize();
        if (value && !d.trace) {
            const size_t bytes = (size_t)d.sm_count * lb2::kTraceItems * lb2::kTraceEvents * sizeof(unsigned long long);
            if (cudaMalloc(&d.trace, bytes) != cudaSuccess) return fail(LB2_ERR_NOMEM, "trace buffer");
            cudaMemset(d.trace, 0, bytes);
        } else if (!value && d.trace) {
            cudaFree(d.trace);
            d.trace = nullptr;
        }
        ctx->option_epoch++;
        return LB2_OK;
    }
    if (!strcmp(name, "profile_reserve")) {
        // pre-create `value` (start, stop) event pairs per device for profile_trunk
        for (auto& dp : ctx->dev) {
            std::lock_guard<std::mutex> lk(dp->mu);
            CU_TRY(cudaSetDevice(dp->id));
            while ((long)dp->prof_pool.size() < 2 * value) {
                cudaEvent_t e;
                CU_TRY(cudaEventCreate(&e));
                dp->prof_pool.push_back(e);
            }
        }
        return LB2_OK;
    }
    std::lock_guard<std::mutex> lk(ctx->opt_mu);
    if (!strcmp(name, "trunk_mode")) {
        if (value != 0 && value != 1) return fail(LB2_ERR_INVALID, "trunk_mode must be 0 or 1");
        ctx->trunk_mode = value;
    } else if (!strcmp(name, "cta_pair")) {
        ctx->cta_pair = value ? 1 : 0;
    } else if (!strcmp(name, "dynamic_items")) {
        ctx->dynamic_items = value ? 1 : 0;
    } else if (!strcmp(name, "use_graphs")) {
        ctx->use_graphs = value ? 1 : 0;
    } else if (!strcmp(name, "spin_wait")) {
        ctx->spin_wait = value ? 1 : 0;
    } else if (!strcmp(name, "queue_linger")) {
        ctx->queue_linger = value ? 1 : 0;
    } else if (!strcmp(name, "small_batch")) {
        if (value < 0) return fail(LB2_ERR_INVALID, "small_batch must be >= 0");
        ctx->small_batch = value;
    } else if (!strcmp(name, "group_positions")) {
        if (value < 0 || (value % 128)) return fail(LB2_ERR_INVALID, "group_positions must be 0 or a multiple of 128");
        ctx->group_positions = value;
    } else if (!strcmp(name, "resident_weights")) {
        if (value < 0 || value > 2) return fail(LB2_ERR_INVALID, "resident_weights must be 0, 1 or 2");
        ctx->resident_weights = value;
    } else if (!strcmp(name, "precise")) {
        // 1: both nets in full split-operand precision; 0: back to the default (policy fp16, value lite)
        ctx->precision[0] = value ? kPrecFull : kPrecFp16;
        ctx->precision[1] = value ? kPrecFull : kPrecLite;
    } else if (!strcmp(name, "policy_precision") || !strcmp(name, "value_precision")) {
        if (value < kPrecFp16 || value > kPrecFull) return fail(LB2_ERR_INVALID, "%s must be 0 (fp16), 1 (lite) or 2 (full)", name);
        const int k = name[0] == 'p' ? 0 : 1;
        if (value != kPrecFp16 && ctx->precision[1 - k] != kPrecFp16 && ctx->precision[1 - k] != value)
            return fail(LB2_ERR_UNSUPPORTED, "lite and full precision cannot be mixed between the nets");
        ctx->precision[k] = value;
    } else if (!strcmp(name, "policy_clusters")) {
        ctx->policy_clusters = value;
    } else if (!strcmp(name, "profile_trunk")) {
        ctx->profile_trunk = value ? 1 : 0;
    } else if (!strcmp(name, "max_batch")) {
        if (value < 1 || value > 65536) return fail(LB2_ERR_INVALID, "max_batch out of range");
        ctx->max_batch = value;
    } else {
        return fail(LB2_ERR_INVALID, "unknown option %s", name);
    }
    ctx->option_epoch++;
    return LB2_OK;
}

long lb2_get_option(lb2_ctx* ctx, const char* name) {
    if (!ctx || !name) return -1;
    if (!strcmp(name, "sm_count")) return ctx->dev.empty() ? 0 : ctx->dev[0]->sm_count;
    if (!strcmp(name, "stat_positions")) return ctx->stat_positions.load();
    if (!strcmp(name, "stat_batches")) return ctx->stat_batches.load();
    if (!strcmp(name, "stat_requests")) return ctx->stat_requests.load();
    if (!strcmp(name, "graph_launches")) return ctx->graph_launches.load();
    if (!strcmp(name, "trunk_ns") || !strcmp(name, "trunk_launches_timed")) {
        // device time spent in trunk launches since the last query (profile_trunk = 1); the events go back to the pool
        double ms_total = 0;
        long pairs = 0;
        for (auto& dp : ctx->dev) {
            DeviceState& d = *dp;
            std::lock_guard<std::mutex> lk(d.mu);
            cudaSetDevice(d.id);
            for (size_t i = 0; i + 1 < d.prof_events.size(); i += 2) {
                float ms = 0;
                cudaEventSynchronize(d.prof_events[i + 1]);
                if (cudaEventElapsedTime(&ms, d.prof_events[i], d.prof_events[i + 1]) == cudaSuccess) ms_total += ms;
                else cudaGetLastError();
                d.prof_pool.push_back(d.prof_events[i]);
                d.prof_pool.push_back(d.prof_events[i + 1]);
                pairs++;
            }
            d.prof_events.clear();
        }
        return !strcmp(name, "trunk_ns") ? (long)(ms_total * 1e6) : pairs;
    }
    std::lock_guard<std::mutex> lk(ctx->opt_mu);
    if (!strcmp(name, "trunk_mode")) return ctx->trunk_mode;
    if (!strcmp(name, "max_batch")) return ctx->max_batch;
    if (!strcmp(name, "cta_pair")) return ctx->cta_pair;
    if (!strcmp(name, "dynamic_items")) return ctx->dynamic_items;
    if (!strcmp(name, "use_graphs")) return ctx->use_graphs;
    if (!strcmp(name, "spin_wait")) return ctx->spin_wait;
    if (!strcmp(name, "queue_linger")) return ctx->queue_linger;
    if (!strcmp(name, "group_positions")) return ctx->group_positions;
    if (!strcmp(name, "small_batch")) return ctx->small_batch;
    if (!strcmp(name, "resident_weights")) return ctx->resident_weights;
    if (!strcmp(name, "precise")) return ctx->precision[0] == kPrecFull && ctx->precision[1] == kPrecFull;
    if (!strcmp(name, "policy_precision")) return ctx->precision[0];
    if (!strcmp(name, "value_precision")) return ctx->precision[1];
    if (!strcmp(name, "policy_clusters")) return ctx->policy_clusters;
    if (!strcmp(name, "profile_trunk")) return ctx->profile_trunk;
    return -1;
}

long lb2_launch_count(lb2_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }

int lb2_debug_read_trace(lb2_ctx* ctx, unsigned long long* out, long max_entries) {
    if (!ctx || !out) return fail(LB2_ERR_INVALID, "null argument");
    DeviceState& d = *ctx->dev[0];
    std::lock_guard<std::mutex> lk(d.mu);
    if (!d.trace) return fail(LB2_ERR_STATE, "tracing not enabled");
    const long n = std::min<long>(max_entries, (long)d.sm_count * lb2::kTraceItems * lb2::kTraceEvents);
    CU_TRY(cudaSetDevice(d.id));
    CU_TRY(cudaDeviceSynchronize());
    CU_TRY(cudaMemcpy(out, d.trace, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemset(d.trace, 0, (size_t)n * sizeof(unsigned long long)));
    return (int)(n / (lb2::kTraceItems * lb2::kTraceEvents));
}

int lb2_debug_trunk(lb2_ctx* ctx, int kind, const uint32_t* planes, const uint8_t* rotation, int n, int n_layers,
                    float* act_out) {
    if (kind != LB2_POLICY && kind != LB2_VALUE) return fail(LB2_ERR_INVALID, "bad kind");
    bool need[2] = {kind == LB2_POLICY, kind == LB2_VALUE};
    int rc = check_ready(ctx, need);
    if (rc) return rc;
    if (n <= 0 || !planes || !rotation || !act_out) return fail(LB2_ERR_INVALID, "bad arguments");
    if ((rc = check_rotations(rotation, n))) return rc;
    Options o = snapshot(ctx);
    o.profile_trunk = 0;
    DeviceState* d = ctx->dev[0].get();
    std::lock_guard<std::mutex> lk(d->mu);
    CU_TRY(cudaSetDevice(d->id));
    NetDev& nd = d->net[kind];
    if (n_layers < 1 || n_layers > (int)nd.trunk.size()) return fail(LB2_ERR_INVALID, "n_layers out of range");
    if (d->cap < n) {
        CU_TRY(cudaDeviceSynchronize());
        cudaFree(d->rot);
        d->rot = nullptr;
        d->cap = 0;
        CU_TRY(cudaMalloc(&d->rot, n));
        d->cap = n;
    }
    if ((rc = grow_workspaces(d, need, n))) return rc;
    // the inputs go up on the compute stream: behind every earlier evaluation that read nd.planes / d->rot
    if (d->last_on_user_stream) CU_TRY(cudaStreamWaitEvent(d->stream, d->ev_last, 0));
    CU_TRY(cudaMemcpyAsync(d->rot, rotation, n, cudaMemcpyHostToDevice, d->stream));
    CU_TRY(cudaMemcpyAsync(nd.planes, planes, (size_t)n * lb2::kPoints * sizeof(uint32_t), cudaMemcpyHostToDevice, d->stream));
    EvalKey key;
    key.n = n;
    key.limit[kind] = n_layers;
    key.in[kind] = nd.planes;
    key.in[2] = d->rot;
    o.use_graphs = 0;
    if ((rc = eval_on_device(ctx, o, d, key, -2, d->stream))) return rc;
    const __half* last = nd.act[(n_layers - 1) & 1];
    const int c_out = nd.trunk[n_layers - 1].c_out;
    std::vector<__half> host((size_t)(c_out / 8) * nd.rows3 * 8);
    CU_TRY(cudaMemcpyAsync(host.data(), last, host.size() * sizeof(__half), cudaMemcpyDeviceToHost, d->stream));
    CU_TRY(cudaStreamSynchronize(d->stream));
    for (int p = 0; p < n; p++)
        for (int c = 0; c < c_out; c++)
            for (int y = 0; y < 19; y++)
                for (int x = 0; x < 19; x++) {
                    const size_t row = (size_t)p * 400 + y * 20 + x;
                    act_out[((size_t)p * c_out + c) * 361 + y * 19 + x] =
                        __half2float(host[((size_t)(c / 8) * nd.rows3 + row) * 8 + (c % 8)]);
                }
    return LB2_OK;
}

}  // extern "C"
