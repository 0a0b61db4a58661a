// lb2_api.cu — host side of the C ABI declared in include/leela_b200.h.
// Owns devices, replicated weights, per-device workspaces, TMA tensor maps and the batch
// pipeline: pinned staging -> H2D -> expand -> trunk (tcgen05) -> heads -> D2H.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
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

// One trunk layer on a device. A layer wider than 128 output channels is stored (and later run) as
// n_split column splits of c_out / n_split channels each, every split packed like a layer of its own.
struct TrunkLayerDev {
    int k, c_in, c_out;
    int n_split = 1;
    __half* wpk[lb2::kMaxSplit] = {nullptr, nullptr};
    __half* wpk2[lb2::kMaxSplit] = {nullptr, nullptr};  // CTA-pair packing
    __half* wpk_x[lb2::kMaxSplit] = {nullptr, nullptr};   // the same two packings for precise mode (virtual slabs, see
    __half* wpk2_x[lb2::kMaxSplit] = {nullptr, nullptr};  // pack_trunk_weights)
    float* bias[lb2::kMaxSplit] = {nullptr, nullptr};
};

// Workspace sets: 0 = calls on device pointers (lb2_eval_both_device, lb2_debug_trunk), 1 + i = I/O slot i.
constexpr int kIoSlots = 2;
constexpr int kSets = 1 + kIoSlots;

// One net replicated on one device, with its workspace.
struct NetDev {
    std::vector<TrunkLayerDev> trunk;
    int head_c_in = 0;
    float *head_wt[lb2::kMaxSplit] = {nullptr, nullptr};  // final conv weights per column split of the last trunk layer, [9 taps][c_in / n_split]
    float* head_b = nullptr;
    // Buffers written OUTSIDE the trunk launch exist once per workspace set (kSets below), so that
    // the expand and heads kernels of one call can overlap the trunk of another:
    float* zbuf[1 + kIoSlots] = {};               // fused-head partial sums [splits * parts][9][rows3]
    int hidden = 0;
    float *ip1_wt = nullptr, *ip1_b = nullptr, *ip2_w = nullptr, *ip2_b = nullptr;
    // workspace
    int cap = 0;
    int width = 0;           // widest trunk c_out
    int rows5 = 0, rows3 = 0;  // chunk-plane rows of the S=21 / S=20 buffers
    uint32_t* planes = nullptr;
    __half* x0[1 + kIoSlots] = {};                // first-conv input (S=21 row space)
    __half* act[2] = {nullptr, nullptr};
    uint32_t* flags = nullptr;
    int flags_stride = 0;
    CUtensorMap tm_x0[1 + kIoSlots], tm_act[2];
};

// Input/output buffers of one host-buffer call in flight on a device. Two slots per device let the
// copies of one call overlap the kernels of another (the activation workspace is shared, so the
// kernels themselves run one call after the other on the device's compute stream).
struct IoSlot {
    cudaStream_t stream = nullptr;          // copies of this slot
    cudaEvent_t ev_in = nullptr, ev_done = nullptr;
    int cap = 0;
    bool busy = false;
    uint32_t* d_planes[2] = {nullptr, nullptr};
    uint8_t* d_rot = nullptr;
    float *d_probs = nullptr, *d_win = nullptr;
    uint32_t* h_planes[2] = {nullptr, nullptr};  // pinned staging for pageable caller buffers
    uint8_t* h_rot = nullptr;
    float *h_probs = nullptr, *h_win = nullptr;
};
struct DeviceState {
    int id = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    NetDev net[2];
    uint8_t* rot = nullptr;               // rotation buffer of lb2_debug_trunk
    lb2::LayerJob* jobs_dev[kSets] = {};   // one job table per workspace set (they differ in x0 / zbuf)
    uint32_t* item_counter = nullptr;     // dynamic scheduling: claim counter, never reset
    uint32_t claim_base = 0;              // its value at the start of the next trunk launch
    uint32_t net_claim_base[2] = {0, 0};  // the same for the per-net counters of the resident-weights mode
    lb2::LayerJob* h_jobs[kSets] = {};    // pinned staging of the job tables
    int cap = 0;                          // capacity of `rot`
    uint32_t epoch = 0;
    unsigned long long* trace = nullptr;   // debug timeline buffer (option "trace")
    std::vector<cudaEvent_t> prof_events;  // (start, stop) pairs around trunk launches
    std::vector<cudaEvent_t> seg_events;   // profile_trunk == 2: one event around every launch, 4 per eval
    long plan_key[kSets][8];  // n, run0, run1, limit0, limit1, workspace pointers, pair (-1 = none yet)
    IoSlot slots[kIoSlots];
    cudaEvent_t ev_user = nullptr;   // last work enqueued on a caller-provided stream (lb2_eval_both_device)
    cudaEvent_t ev_comp = nullptr;   // marks the compute stream for such a call to wait on
    bool user_pending = false;
};

struct Request {
    int kind;
    std::vector<uint32_t> planes;
    std::vector<uint8_t> rot;
    int n;
    float temp;
    float* out;
    lb2_callback cb;
    void* user;
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
    std::vector<DeviceState> dev;
    std::unique_ptr<lb2_net> nets[2];
    std::string backend;
    long trunk_mode = 1;
    long max_batch = 256;  // batch-256 chunks keep both nets' ping-pong activations L2-resident (measured best)
    long profile_trunk = 0;
    long cta_pair = 1;
    long overlap_io = 0;   // 1: host-buffer calls run expand / heads on the I/O slot's stream, beside the trunk of another
                           // call. Measured 2-3 % slower end to end (the value head's blocks hold up the next trunk's CTAs): off.
    long dynamic_items = 1;
    long resident_weights = LB2_RESIDENT_DEFAULT;   // keep each CTA's half of a layer's weights in shared memory across the layer's items
    long precise = 0;      // split-operand mode: activations and weights as fp16 hi + fp16 lo, three MMA terms (hi*Wh + hi*Wl + lo*Wh)
                           // accumulated in fp32 -> fp32-grade results at about a third of the throughput
    long policy_clusters = -1;   // resident mode: clusters that prefer the policy net (-1 = split by estimated work)
    std::atomic<long> launches{0};
    std::atomic<long> stat_positions{0}, stat_batches{0}, stat_requests{0};  // async queue: positions, device batches, requests
    std::mutex eval_mu;                 // guards enqueueing on the devices and the shared workspaces
    std::mutex slot_mu;                 // guards IoSlot::busy
    std::condition_variable slot_cv;
    // async submission
    std::mutex q_mu;
    std::condition_variable q_cv, q_idle;
    std::deque<Request> queue;
    bool worker_run = false;
    int workers_busy = 0;
    std::vector<std::thread> workers;   // two: one batch's copies overlap the other's kernels
};

namespace {

// --------------------------------------------------------------------------------------------
// weights
// --------------------------------------------------------------------------------------------
void tap_groups(int k, int* n, int* begin, int* end) {
    if (k == 3) { *n = 1; begin[0] = 0; end[0] = 9; }
    else { *n = 3; begin[0] = 0; end[0] = 9; begin[1] = 9; end[1] = 17; begin[2] = 17; end[2] = 25; }
}

// Virtual K slabs of a layer. Plain mode: slab s = channels [16 s, 16 s + 16) of the fp16-rounded weights.
// Precise mode (`terms` = 3, or 2 for the first layer whose binary inputs have no residual): fp32 weight
// w = Wh + Wl (+ 2^-22 w) with Wh = fp16(w), Wl = fp16(w - Wh); activation a = hi + lo likewise. The K loop
// runs over [hi x Wh | hi x Wl | lo x Wh], i.e. virtual slab v takes weights slab v % n from Wh (v < n or
// v >= 2n) or Wl (n <= v < 2n); the kernel's producer maps v to the matching activation chunk planes.
__half slab_weight(const HostConv& c, int v, int n_real, int co, int j, int e, int t) {
    const int kk = c.k * c.k;
    const int ci = 16 * (v % n_real) + 8 * j + e;
    const float w = c.w[((size_t)co * c.c_in + ci) * kk + t];
    const __half wh = __float2half_rn(w);
    if (v >= n_real && v < 2 * n_real) return __float2half_rn(w - __half2float(wh));
    return wh;
}

// [slab][tap group][tap][2 chunks][c_out][8] fp16 — the smem image of each pipeline stage.
std::vector<__half> pack_trunk_weights(const HostConv& c, int terms = 1) {
    const int kk = c.k * c.k, n_real = c.c_in / 16;
    std::vector<__half> out((size_t)kk * c.c_in * c.c_out * terms);
    int ng, gb[3], ge[3];
    tap_groups(c.k, &ng, gb, ge);
    size_t o = 0;
    for (int s = 0; s < n_real * terms; s++)
        for (int g = 0; g < ng; g++)
            for (int t = gb[g]; t < ge[g]; t++)
                for (int j = 0; j < 2; j++)
                    for (int n = 0; n < c.c_out; n++)
                        for (int e = 0; e < 8; e++) out[o++] = slab_weight(c, s, n_real, n, j, e, t);
    return out;
}

// CTA-pair packing: [slab][tap group][rank][tap][2 chunks][c_out/2][8]; rank r holds output channels
// [r*c_out/2, (r+1)*c_out/2) — the half of the MMA's B operand that CTA r of the pair stages.
std::vector<__half> pack_trunk_weights_pair(const HostConv& c, int terms = 1) {
    const int kk = c.k * c.k, nh = c.c_out / 2, n_real = c.c_in / 16;
    std::vector<__half> out((size_t)kk * c.c_in * c.c_out * terms);
    int ng, gb[3], ge[3];
    tap_groups(c.k, &ng, gb, ge);
    size_t o = 0;
    for (int s = 0; s < n_real * terms; s++)
        for (int g = 0; g < ng; g++)
            for (int r = 0; r < 2; r++)
                for (int t = gb[g]; t < ge[g]; t++)
                    for (int j = 0; j < 2; j++)
                        for (int n = 0; n < nh; n++)
                            for (int e = 0; e < 8; e++) out[o++] = slab_weight(c, s, n_real, r * nh + n, j, e, t);
    return out;
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
        t.n_split = n_splits(c.c_out);
        const int w = c.c_out / t.n_split;
        for (int sp = 0; sp < t.n_split; sp++) {
            HostConv part;   // output channels [sp*w, (sp+1)*w): a contiguous block of the OIHW array
            part.k = c.k; part.c_in = c.c_in; part.c_out = w;
            const size_t per_out = (size_t)c.c_in * c.k * c.k;
            part.w.assign(c.w.begin() + (size_t)sp * w * per_out, c.w.begin() + (size_t)(sp + 1) * w * per_out);
            part.b.assign(c.b.begin() + sp * w, c.b.begin() + (sp + 1) * w);
            std::vector<__half> pk = pack_trunk_weights(part);
            int rc = upload(&t.wpk[sp], pk.data(), pk.size() * sizeof(__half));
            if (rc) return rc;
            std::vector<__half> pk2 = pack_trunk_weights_pair(part);
            rc = upload(&t.wpk2[sp], pk2.data(), pk2.size() * sizeof(__half));
            if (rc) return rc;
            const int terms = (l == 0) ? 2 : 3;   // precise mode; the first layer's inputs are 0/1: no residual term
            pk = pack_trunk_weights(part, terms);
            if ((rc = upload(&t.wpk_x[sp], pk.data(), pk.size() * sizeof(__half)))) return rc;
            pk2 = pack_trunk_weights_pair(part, terms);
            if ((rc = upload(&t.wpk2_x[sp], pk2.data(), pk2.size() * sizeof(__half)))) return rc;
            rc = upload(&t.bias[sp], part.b.data(), part.b.size() * sizeof(float));
            if (rc) return rc;
        }
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
    cudaFree(nd->flags);
    for (int w = 0; w < kSets; w++) { cudaFree(nd->x0[w]); cudaFree(nd->zbuf[w]); nd->x0[w] = nullptr; nd->zbuf[w] = nullptr; }
    nd->planes = nullptr; nd->act[0] = nd->act[1] = nullptr; nd->flags = nullptr;
    nd->cap = 0;
}

int ensure_workspace(NetDev* nd, int kind, int cap) {
    if (nd->cap >= cap) return LB2_OK;
    free_workspace(nd);
    nd->rows5 = round_up(cap * 441, 2 * lb2::kTileRows);  // whole CTA-pair items
    nd->rows3 = round_up(cap * 400, 2 * lb2::kTileRows);
    const size_t x0_bytes = (size_t)4 * nd->rows5 * 16;
    const size_t act_bytes = (size_t)(2 * nd->width / 8) * nd->rows3 * 16;   // hi planes + (precise mode) lo planes
    CU_TRY(cudaMalloc(&nd->planes, (size_t)cap * lb2::kPoints * sizeof(uint32_t)));
    for (int w = 0; w < kSets; w++) {
        CU_TRY(cudaMalloc(&nd->x0[w], x0_bytes));
        CU_TRY(cudaMemset(nd->x0[w], 0, x0_bytes));
        CU_TRY(cudaMalloc(&nd->zbuf[w], (size_t)9 * lb2::kColParts * lb2::kMaxSplit * nd->rows3 * sizeof(float)));
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
    for (int w = 0; w < kSets; w++)
        if ((rc = make_act_tmap(&nd->tm_x0[w], nd->x0[w], nd->rows5, 4, 48))) return rc;
    if ((rc = make_act_tmap(&nd->tm_act[0], nd->act[0], nd->rows3, 2 * nd->width / 8, 24))) return rc;
    if ((rc = make_act_tmap(&nd->tm_act[1], nd->act[1], nd->rows3, 2 * nd->width / 8, 24))) return rc;
    nd->cap = cap;
    return LB2_OK;
}

// device-side rotation buffer of lb2_debug_trunk (host-buffer calls use their I/O slot's)
int ensure_device_staging(DeviceState* d, int cap) {
    if (d->cap >= cap) return LB2_OK;
    cudaFree(d->rot);
    d->rot = nullptr;
    CU_TRY(cudaMalloc(&d->rot, cap));
    d->cap = cap;
    return LB2_OK;
}

// --------------------------------------------------------------------------------------------
// the batch pipeline on one device; all pointers are device pointers
// --------------------------------------------------------------------------------------------
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
JobPlan plan_jobs(DeviceState* d, bool run[2], int n, int limit_layers[2], bool pair, int ws, bool net_major, bool precise) {
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
            const int w = t.c_out / t.n_split;
            for (int sp = 0; sp < t.n_split; sp++) {
                lb2::LayerJob J;
                memset(&J, 0, sizeof J);
                const bool first = (l == 0);
                J.S = first ? 21 : 20;
                J.ksize = t.k;
                J.halo = first ? 48 : 24;
                J.n_real_slabs = t.c_in / 16;
                J.n_slabs = precise ? (first ? 2 : 3) * J.n_real_slabs : J.n_real_slabs;
                J.lo_chunks = precise ? t.c_out / 8 : 0;   // the lo planes of a layer's output follow its c_out / 8 hi planes
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
                J.n_pos = n;
                J.net = k;
                J.layer = (int)l;
                J.wpk = precise ? t.wpk_x[sp] : t.wpk[sp];
                J.wpk2 = precise ? t.wpk2_x[sp] : t.wpk2[sp];
                J.bias = t.bias[sp];
                J.flags = nd.flags + ((size_t)l * lb2::kMaxSplit + sp) * nd.flags_stride;
                J.head_slot = net_major ? k : k * lb2::kMaxSplit + sp;
                if (l + 1 == nd.trunk.size() && limit_layers[k] > (int)nd.trunk.size()) {
                    // whole net: fold the final 3x3 conv to one channel into this layer's epilogue
                    J.head_taps = 9;
                    J.head_w = nd.head_wt[sp];
                    J.zbuf = nd.zbuf[ws];
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
            prev_split[k] = t.n_split;
        }
    }
    return pl;
}

// `ws`: workspace set (x0 / zbuf / job table) this launch works in
int run_trunk(lb2_ctx* ctx, DeviceState* d, bool run[2], int n, int limit_layers[2], cudaStream_t st,
              JobPlan* plan_out, int ws) {
    const bool pair = ctx->cta_pair != 0 && d->sm_count >= 2;
    // resident-weights mode: CTA pairs, the single dataflow launch with dynamic claiming, and every layer's
    // half of the packed weights must fit the resident area (c_in, c_out <= 128: no column splits)
    const bool precise = ctx->precise != 0;
    const bool want_resident = ctx->resident_weights == 1 || (ctx->resident_weights == 2 && run[0] && run[1]);
    bool resident = pair && !precise && want_resident && ctx->trunk_mode == 1 && ctx->dynamic_items != 0;
    int n_jobs_est = 0;
    for (int k = 0; k < 2 && resident; k++) {
        if (!run[k]) continue;
        for (size_t l = 0; l < d->net[k].trunk.size() && (int)l < limit_layers[k]; l++) {
            const TrunkLayerDev& t = d->net[k].trunk[l];
            if (t.n_split != 1 || (size_t)t.k * t.k * t.c_in * (t.c_out / 2) * 2 > (size_t)lb2::kResWeightBytes) resident = false;
            n_jobs_est++;
        }
    }
    if (n_jobs_est > lb2::kResJobs) resident = false;
    JobPlan pl = plan_jobs(d, run, n, limit_layers, pair, ws, resident, precise);
    if (pl.jobs.empty()) return fail(LB2_ERR_STATE, "no trunk layers to run");
    if ((int)pl.jobs.size() > lb2::kMaxLaunchJobs) return fail(LB2_ERR_UNSUPPORTED, "too many layers");
    const long key[8] = {n, run[0], run[1], limit_layers[0], limit_layers[1],
                         (long)reinterpret_cast<uintptr_t>(d->net[0].act[0]), (long)reinterpret_cast<uintptr_t>(d->net[1].act[0]),
                         (long)pair + 2 * (long)resident + 4 * (long)precise};
    if (memcmp(key, d->plan_key[ws], sizeof key)) {
        // the set's job table on the device is reused by back-to-back launches of the same shape;
        // rewrite it only when the shape changes, after earlier work has drained
        CU_TRY(cudaStreamSynchronize(st));
        CU_TRY(cudaStreamSynchronize(d->stream));
        memcpy(d->h_jobs[ws], pl.jobs.data(), pl.jobs.size() * sizeof(lb2::LayerJob));
        CU_TRY(cudaMemcpyAsync(d->jobs_dev[ws], d->h_jobs[ws], pl.jobs.size() * sizeof(lb2::LayerJob), cudaMemcpyHostToDevice, st));
        CU_TRY(cudaStreamSynchronize(st));
        memcpy(d->plan_key[ws], key, sizeof key);
    }
    lb2::TrunkParams P;
    memset(&P, 0, sizeof P);
    for (int k = 0; k < 2; k++) {
        if (!d->net[k].cap) continue;
        P.tmaps[pl.tmap_base[k] + 0] = d->net[k].tm_x0[ws];
        P.tmaps[pl.tmap_base[k] + 1] = d->net[k].tm_act[0];
        P.tmaps[pl.tmap_base[k] + 2] = d->net[k].tm_act[1];
    }
    for (int k = 0; k < 2; k++)  // unused slots still get prefetched: point them at a valid map
        if (!d->net[k].cap)
            for (int i = 0; i < 3; i++) P.tmaps[pl.tmap_base[k] + i] = d->net[1 - k].tm_x0[ws];
    P.jobs = d->jobs_dev[ws];
    P.n_jobs = (int)pl.jobs.size();
    // rounds: jobs of equal depth, their items interleaved in the launch-wide order
    P.n_rounds = 0;
    P.round_base[0] = 0;
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
    for (int r = P.n_rounds + 1; r <= lb2::kMaxRounds; r++) P.round_base[r] = 0x7fffffff;
    P.epoch = ++d->epoch;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (ctx->profile_trunk) {
        CU_TRY(cudaEventCreate(&ev0));
        CU_TRY(cudaEventCreate(&ev1));
        CU_TRY(cudaEventRecord(ev0, st));
    }
    if (const char* dbg = getenv("LB2_DEBUG_FLAGS")) P.debug_flags = atoi(dbg);
    P.trace = d->trace;
    if (ctx->trunk_mode == 1) {
        P.item_begin = 0;
        P.item_end = pl.total_items;
        P.use_flags = 1;
        P.next_item = ctx->dynamic_items ? d->item_counter : nullptr;
        int grid = pair ? std::min(d->sm_count & ~1, 2 * pl.total_items) : std::min(d->sm_count, pl.total_items);
        if (const char* g = getenv("LB2_GRID")) grid = std::max(2, std::min(grid, atoi(g) & ~1));   // experiment: fewer SMs (power-cap study)
        const int n_clusters = pair ? grid / 2 : grid;
        if (resident) {
            // each cluster draws items of its preferred net until that counter runs past the end, then helps the
            // other net: one end marker per cluster per net, i.e. items + clusters claims per net per launch
            for (int k = 0; k < 2; k++) {
                P.net_next_item[k] = d->item_counter + 1 + k;
                P.net_claim_base[k] = d->net_claim_base[k];
                P.net_item_begin[k] = pl.net_begin[k];
                P.net_item_end[k] = pl.net_end[k];
                d->net_claim_base[k] += (uint32_t)(pl.net_end[k] - pl.net_begin[k]) + (uint32_t)n_clusters;
            }
            const double total = pl.net_cost[0] + pl.net_cost[1];
            int pc = total > 0 ? (int)(n_clusters * pl.net_cost[0] / total + 0.5) : n_clusters;
            if (pl.net_cost[0] > 0 && pl.net_cost[1] > 0) pc = std::max(1, std::min(n_clusters - 1, pc));
            if (ctx->policy_clusters >= 0 && pl.net_cost[0] > 0 && pl.net_cost[1] > 0) pc = (int)std::min<long>(n_clusters - 1, std::max<long>(1, ctx->policy_clusters));
            P.policy_clusters = pc;
        } else if (P.next_item) {
            // every cluster claims until it draws an index past the end: items + clusters claims per launch
            P.claim_base = d->claim_base;
            d->claim_base += (uint32_t)pl.total_items + (uint32_t)n_clusters;
        }
        CU_TRY(lb2::launch_trunk(P, grid, true, pair, resident, precise, st));
        ctx->launches++;
    } else {
        P.use_flags = 0;
        for (int r = 0; r < P.n_rounds; r++) {
            P.item_begin = P.round_base[r];
            P.item_end = P.round_base[r + 1];
            const int items = P.item_end - P.item_begin;
            const int grid = pair ? std::min(d->sm_count & ~1, 2 * items) : std::min(d->sm_count, items);
            CU_TRY(lb2::launch_trunk(P, grid, false, pair, false, precise, st));
            ctx->launches++;
        }
    }
    if (ev0) {
        CU_TRY(cudaEventRecord(ev1, st));
        d->prof_events.push_back(ev0);
        d->prof_events.push_back(ev1);
    }
    if (plan_out) *plan_out = pl;
    return LB2_OK;
}

// One evaluation on one device; all pointers are device pointers. The expand and heads kernels go on
// `st_io`, the trunk on `st`; when the two differ (host-buffer calls: the I/O slot's stream and the
// device's compute stream) `ev_in` / `ev_done` order them, and the expand / heads of one call run
// beside the trunk of another. `ws` is the workspace set (x0, zbuf, job table) the call works in.
// `ensemble`: n device positions = n/8 input positions x 8 symmetries (AVERAGE_ALL), d_rot unused.
int eval_on_device(lb2_ctx* ctx, DeviceState* d, const uint32_t* d_pol, const uint32_t* d_val, const uint8_t* d_rot,
                   int n, float temp, float* d_probs, float* d_win, cudaStream_t st, int ws = 0, bool ensemble = false,
                   cudaStream_t st_io = nullptr, cudaEvent_t ev_in = nullptr, cudaEvent_t ev_done = nullptr, bool overlap = true) {
    bool run[2] = {d_probs != nullptr, d_win != nullptr};
    const bool split = st_io != nullptr && st_io != st;
    if (!split) st_io = st;
    // overlap off: the expand and heads kernels go on the compute stream too (the two streams are still ordered)
    cudaStream_t st_k = overlap ? st_io : st;
    // profile_trunk == 2: an event after every launch -> per-segment device times (debug)
    auto mark = [&](cudaStream_t s) {
        if (ctx->profile_trunk != 2) return;
        cudaEvent_t e;
        if (cudaEventCreate(&e) == cudaSuccess) { cudaEventRecord(e, s); d->seg_events.push_back(e); }
    };
    if (split && !overlap) {   // inputs arrive on st_io
        CU_TRY(cudaEventRecord(ev_in, st_io));
        CU_TRY(cudaStreamWaitEvent(st, ev_in, 0));
    }
    mark(st_k);
    const uint32_t* planes[2] = {d_pol, d_val};
    int limit[2] = {1 << 20, 1 << 20};
    lb2::ExpandArgs ea;
    memset(&ea, 0, sizeof ea);
    ea.rotation = d_rot;
    ea.ensemble = ensemble ? 1 : 0;
    ea.n = n;
    for (int k = 0; k < 2; k++) {
        if (!run[k]) continue;
        NetDev& nd = d->net[k];
        if (nd.cap < n) return fail(LB2_ERR_STATE, "workspace too small");
        ea.planes[ea.n_nets] = planes[k];
        ea.x0[ea.n_nets] = nd.x0[ws];
        ea.chunk_rows[ea.n_nets] = nd.rows5;
        ea.n_nets++;
        if (k == 1 && nd.ip1_wt) {  // pull the value head's matrix into L2 while the trunk runs
            ea.pf = reinterpret_cast<const uint8_t*>(nd.ip1_wt);
            ea.pf_bytes = (size_t)lb2::kPoints * nd.hidden * sizeof(float);
        }
    }
    CU_TRY(lb2::launch_expand(ea, st_k));
    ctx->launches++;
    mark(st_k);
    if (split && overlap) {
        CU_TRY(cudaEventRecord(ev_in, st_io));
        CU_TRY(cudaStreamWaitEvent(st, ev_in, 0));
    }
    JobPlan pl;
    int rc = run_trunk(ctx, d, run, n, limit, st, &pl, ws);
    if (rc) return rc;
    mark(st);
    if (split && overlap) {
        CU_TRY(cudaEventRecord(ev_done, st));
        CU_TRY(cudaStreamWaitEvent(st_io, ev_done, 0));
    }
    lb2::HeadArgs ha;
    memset(&ha, 0, sizeof ha);
    ha.rotation = d_rot;
    ha.ensemble = ensemble ? 1 : 0;
    ha.temp = temp;
    if (run[0]) {
        NetDev& nd = d->net[0];
        ha.p_zbuf = nd.zbuf[ws]; ha.p_chunk_rows = nd.rows3; ha.p_bias = nd.head_b; ha.probs = d_probs; ha.n_policy = n;
        ha.p_parts = nd.trunk.back().n_split * lb2::kColParts;
    }
    if (run[1]) {
        NetDev& nd = d->net[1];
        ha.v_zbuf = nd.zbuf[ws]; ha.v_chunk_rows = nd.rows3; ha.v_bias = nd.head_b; ha.ip1_wt = nd.ip1_wt; ha.ip1_b = nd.ip1_b;
        ha.hidden = nd.hidden; ha.ip2_w = nd.ip2_w; ha.ip2_b = nd.ip2_b; ha.winrate = d_win; ha.n_value = n;
        ha.v_parts = nd.trunk.back().n_split * lb2::kColParts;
    }
    ha.trace = d->trace;
    ha.trace_ctas = d->sm_count;
    CU_TRY(lb2::launch_heads(ha, st_k));
    ctx->launches++;
    mark(st_k);
    if (split && !overlap) {   // results are fetched on st_io
        CU_TRY(cudaEventRecord(ev_done, st));
        CU_TRY(cudaStreamWaitEvent(st_io, ev_done, 0));
    }
    return LB2_OK;
}

int check_ready(lb2_ctx* ctx, bool need[2]) {
    if (!ctx) return fail(LB2_ERR_INVALID, "null context");
    for (int k = 0; k < 2; k++)
        if (need[k] && !(ctx->nets[k] && ctx->nets[k]->finalized))
            return fail(LB2_ERR_STATE, "%s net not finalized", k == 0 ? "policy" : "value");
    return LB2_OK;
}

// true when `p` is page-locked host memory the device can DMA from/to directly
bool is_pinned(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

int check_rotations(const uint8_t* rot, int n) {
    for (int i = 0; i < n; i++)
        if (rot[i] > 7) return fail(LB2_ERR_INVALID, "rotation[%d] = %d out of range 0..7", i, (int)rot[i]);
    return LB2_OK;
}

int ensure_slot(IoSlot* sl, int cap) {
    if (!sl->stream) {
        CU_TRY(cudaStreamCreateWithFlags(&sl->stream, cudaStreamNonBlocking));
        CU_TRY(cudaEventCreateWithFlags(&sl->ev_in, cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&sl->ev_done, cudaEventDisableTiming));
    }
    if (sl->cap >= cap) return LB2_OK;
    CU_TRY(cudaStreamSynchronize(sl->stream));
    for (int k = 0; k < 2; k++) { cudaFree(sl->d_planes[k]); cudaFreeHost(sl->h_planes[k]); }
    cudaFree(sl->d_rot); cudaFree(sl->d_probs); cudaFree(sl->d_win);
    cudaFreeHost(sl->h_rot); cudaFreeHost(sl->h_probs); cudaFreeHost(sl->h_win);
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

void free_slot(IoSlot* sl) {
    for (int k = 0; k < 2; k++) { cudaFree(sl->d_planes[k]); cudaFreeHost(sl->h_planes[k]); }
    cudaFree(sl->d_rot); cudaFree(sl->d_probs); cudaFree(sl->d_win);
    cudaFreeHost(sl->h_rot); cudaFreeHost(sl->h_probs); cudaFreeHost(sl->h_win);
    if (sl->ev_in) cudaEventDestroy(sl->ev_in);
    if (sl->ev_done) cudaEventDestroy(sl->ev_done);
    if (sl->stream) cudaStreamDestroy(sl->stream);
    *sl = IoSlot();
}

// One slot index on EVERY device for the duration of a host-buffer call (so concurrent callers never
// hold slots in opposite orders).
int acquire_slots(lb2_ctx* ctx) {
    std::unique_lock<std::mutex> lk(ctx->slot_mu);
    int si = -1;
    ctx->slot_cv.wait(lk, [&] {
        for (int i = 0; i < kIoSlots; i++)
            if (!ctx->dev[0].slots[i].busy) { si = i; return true; }
        return false;
    });
    for (auto& d : ctx->dev) d.slots[si].busy = true;
    return si;
}
void release_slots(lb2_ctx* ctx, int si) {
    {
        std::lock_guard<std::mutex> lk(ctx->slot_mu);
        for (auto& d : ctx->dev) d.slots[si].busy = false;
    }
    ctx->slot_cv.notify_one();
}

// Host-buffer evaluation: shard positions over devices in contiguous slices, chunk each slice by
// max_batch. Per device the call owns one IoSlot: inputs go up on the slot's stream, the kernels run
// on the device's compute stream (after the previous call's kernels: the activation workspace is
// shared), the results come down on the slot's stream — so with two callers in flight the copies
// of one overlap the kernels of the other. Only the enqueueing is done under the context lock.
int eval_host_locked_enqueue(lb2_ctx* ctx, DeviceState* d, IoSlot* sl, const uint32_t* const src[2], const bool pin_in[2],
                             const uint8_t* rot, bool pin_rot, int lo, int cnt, int cap, float temp, const bool need[2],
                             float* probs, bool pin_probs, float* win, bool pin_win, bool ensemble) {
    int rc;
    CU_TRY(cudaSetDevice(d->id));
    for (int k = 0; k < 2; k++)
        if (need[k] && d->net[k].cap < cap) {
            // growing the shared workspace: nothing (trunk, or another slot's expand / heads) may still be using it
            CU_TRY(cudaDeviceSynchronize());
            if ((rc = ensure_workspace(&d->net[k], k, cap))) return rc;
        }
    const size_t pbytes = (size_t)cnt * lb2::kPoints * sizeof(uint32_t);
    if (!ensemble) CU_TRY(cudaMemcpyAsync(sl->d_rot, pin_rot ? rot + lo : sl->h_rot, cnt, cudaMemcpyHostToDevice, sl->stream));
    for (int k = 0; k < 2; k++) {
        if (!need[k]) continue;
        const uint32_t* from = pin_in[k] ? src[k] + (size_t)lo * lb2::kPoints : sl->h_planes[k];
        CU_TRY(cudaMemcpyAsync(sl->d_planes[k], from, pbytes, cudaMemcpyHostToDevice, sl->stream));
    }
    if (d->user_pending) {   // kernels enqueued on a caller's stream (lb2_eval_both_device) use the same activation buffers
        CU_TRY(cudaStreamWaitEvent(d->stream, d->ev_user, 0));
        d->user_pending = false;
    }
    // ensemble: the 8*cnt per-symmetry results land in the first 8*cnt entries of the slot's output
    // buffers, their means behind them
    const int n_dev = ensemble ? 8 * cnt : cnt;
    rc = eval_on_device(ctx, d, sl->d_planes[0], sl->d_planes[1], sl->d_rot, n_dev, temp, need[0] ? sl->d_probs : nullptr,
                        need[1] ? sl->d_win : nullptr, d->stream, 1 + (int)(sl - d->slots), ensemble, sl->stream, sl->ev_in, sl->ev_done,
                        ctx->overlap_io != 0);
    if (rc) return rc;
    const float* res_probs = sl->d_probs;
    const float* res_win = sl->d_win;
    if (ensemble) {
        lb2::MeanArgs ma;
        memset(&ma, 0, sizeof ma);
        if (need[0]) { ma.probs8 = sl->d_probs; ma.probs = sl->d_probs + (size_t)n_dev * lb2::kPoints; ma.n_policy = cnt; res_probs = ma.probs; }
        if (need[1]) { ma.win8 = sl->d_win; ma.win = sl->d_win + n_dev; ma.n_value = cnt; res_win = ma.win; }
        CU_TRY(lb2::launch_ensemble_mean(ma, sl->stream));
        ctx->launches++;
    }
    if (need[0])
        CU_TRY(cudaMemcpyAsync(pin_probs ? probs + (size_t)lo * lb2::kPoints : sl->h_probs, res_probs,
                               (size_t)cnt * lb2::kPoints * sizeof(float), cudaMemcpyDeviceToHost, sl->stream));
    if (need[1])
        CU_TRY(cudaMemcpyAsync(pin_win ? win + lo : sl->h_win, res_win, (size_t)cnt * sizeof(float), cudaMemcpyDeviceToHost,
                               sl->stream));
    return LB2_OK;
}

int eval_host(lb2_ctx* ctx, const uint32_t* pol, const uint32_t* val, const uint8_t* rot, int n, float temp,
              float* probs, float* win, bool ensemble = false) {
    bool need[2] = {probs != nullptr, win != nullptr};
    int rc = check_ready(ctx, need);
    if (rc) return rc;
    if (n < 0) return fail(LB2_ERR_INVALID, "n < 0");
    if (n == 0) return LB2_OK;
    if ((!ensemble && !rot) || (need[0] && !pol) || (need[1] && !val)) return fail(LB2_ERR_INVALID, "null input pointer");
    if (need[0] && !(temp > 0.0f)) return fail(LB2_ERR_INVALID, "softmax temperature must be > 0");
    if (!ensemble && (rc = check_rotations(rot, n))) return rc;
    // caller buffers that are already page-locked are used directly; pageable ones go through the
    // slot's pinned staging buffers
    const bool pin_in[2] = {is_pinned(pol), is_pinned(val)};
    const bool pin_rot = ensemble || is_pinned(rot), pin_probs = is_pinned(probs), pin_win = is_pinned(win);
    const uint32_t* const src[2] = {pol, val};
    const int ndev = (int)ctx->dev.size();
    const int per = (n + ndev - 1) / ndev;
    long max_batch;
    { std::lock_guard<std::mutex> lk(ctx->eval_mu); max_batch = ctx->max_batch; }
    // an ensemble position occupies 8 device positions (+1 for its mean in the output buffers)
    const int chunk = (int)std::min<long>(ensemble ? std::max<long>(1, max_batch / 8) : max_batch, per);
    const int dev_per_pos = ensemble ? 8 : 1, slot_per_pos = ensemble ? 9 : 1;
    const int si = acquire_slots(ctx);
    struct Release { lb2_ctx* c; int s; ~Release() { release_slots(c, s); } } release{ctx, si};
    for (int base = 0; base < per; base += chunk) {
        std::vector<int> cnt(ndev, 0), off(ndev, 0);
        for (int di = 0; di < ndev; di++) {
            const int lo = std::min(n, di * per + base), hi = std::min(n, std::min((di + 1) * per, di * per + base + chunk));
            cnt[di] = std::max(0, hi - lo);
            off[di] = lo;
            if (!cnt[di]) continue;
            DeviceState* d = &ctx->dev[di];
            IoSlot* sl = &d->slots[si];
            CU_TRY(cudaSetDevice(d->id));
            if ((rc = ensure_slot(sl, slot_per_pos * std::max(chunk, cnt[di])))) return rc;
            // staging copies happen outside the context lock
            if (!pin_rot) memcpy(sl->h_rot, rot + lo, cnt[di]);
            for (int k = 0; k < 2; k++)
                if (need[k] && !pin_in[k])
                    memcpy(sl->h_planes[k], src[k] + (size_t)lo * lb2::kPoints, (size_t)cnt[di] * lb2::kPoints * sizeof(uint32_t));
            std::lock_guard<std::mutex> lk(ctx->eval_mu);
            rc = eval_host_locked_enqueue(ctx, d, sl, src, pin_in, rot, pin_rot, lo, cnt[di], dev_per_pos * std::max(chunk, cnt[di]), temp,
                                          need, probs, pin_probs, win, pin_win, ensemble);
            if (rc) return rc;
        }
        for (int di = 0; di < ndev; di++) {
            if (!cnt[di]) continue;
            DeviceState* d = &ctx->dev[di];
            IoSlot* sl = &d->slots[si];
            CU_TRY(cudaSetDevice(d->id));
            CU_TRY(cudaStreamSynchronize(sl->stream));
            if (need[0] && !pin_probs) memcpy(probs + (size_t)off[di] * lb2::kPoints, sl->h_probs, (size_t)cnt[di] * lb2::kPoints * sizeof(float));
            if (need[1] && !pin_win) memcpy(win + off[di], sl->h_win, (size_t)cnt[di] * sizeof(float));
        }
    }
    return LB2_OK;
}

void worker_loop(lb2_ctx* ctx) {
    for (;;) {
        std::vector<Request> batch;
        {
            std::unique_lock<std::mutex> lk(ctx->q_mu);
            ctx->q_cv.wait(lk, [&] { return !ctx->worker_run || !ctx->queue.empty(); });
            if (!ctx->worker_run && ctx->queue.empty()) return;
            // coalesce every queued request of the same kind and temperature as the head one
            const int kind = ctx->queue.front().kind;
            const float temp = ctx->queue.front().temp;
            int total = 0;
            for (auto it = ctx->queue.begin(); it != ctx->queue.end();) {
                if (it->kind == kind && it->temp == temp && total + it->n <= std::max<long>(ctx->max_batch * (long)ctx->dev.size(), it->n)) {
                    total += it->n;
                    batch.push_back(std::move(*it));
                    it = ctx->queue.erase(it);
                } else {
                    ++it;
                }
            }
            ctx->workers_busy++;
        }
        int total = 0;
        for (auto& r : batch) total += r.n;
        ctx->stat_positions += total; ctx->stat_batches++; ctx->stat_requests += (long)batch.size();
        std::vector<uint32_t> planes((size_t)total * lb2::kPoints);
        std::vector<uint8_t> rot(total);
        int o = 0;
        for (auto& r : batch) {
            memcpy(planes.data() + (size_t)o * lb2::kPoints, r.planes.data(), r.planes.size() * sizeof(uint32_t));
            memcpy(rot.data() + o, r.rot.data(), r.n);
            o += r.n;
        }
        const int kind = batch[0].kind;
        std::vector<float> out(kind == LB2_POLICY ? (size_t)total * lb2::kPoints : (size_t)total);
        int rc = kind == LB2_POLICY
                     ? eval_host(ctx, planes.data(), nullptr, rot.data(), total, batch[0].temp, out.data(), nullptr)
                     : eval_host(ctx, nullptr, planes.data(), rot.data(), total, 1.0f, nullptr, out.data());
        o = 0;
        for (auto& r : batch) {
            if (rc == LB2_OK) {
                const size_t per = kind == LB2_POLICY ? lb2::kPoints : 1;
                memcpy(r.out, out.data() + (size_t)o * per, (size_t)r.n * per * sizeof(float));
            }
            o += r.n;
            if (r.cb) r.cb(r.user, rc);
        }
        {
            std::lock_guard<std::mutex> lk(ctx->q_mu);
            ctx->workers_busy--;
            if (ctx->queue.empty() && ctx->workers_busy == 0) ctx->q_idle.notify_all();
        }
    }
}

int submit(lb2_ctx* ctx, int kind, const uint32_t* planes, const uint8_t* rot, int n, float temp, float* out,
           lb2_callback cb, void* user) {
    bool need[2] = {kind == LB2_POLICY, kind == LB2_VALUE};
    int rc = check_ready(ctx, need);
    if (rc) return rc;
    if (n <= 0 || !planes || !rot || !out) return fail(LB2_ERR_INVALID, "bad submit arguments");
    if ((rc = check_rotations(rot, n))) return rc;
    Request r;
    r.kind = kind; r.n = n; r.temp = temp; r.out = out; r.cb = cb; r.user = user;
    r.planes.assign(planes, planes + (size_t)n * lb2::kPoints);
    r.rot.assign(rot, rot + n);
    {
        std::lock_guard<std::mutex> lk(ctx->q_mu);
        if (!ctx->worker_run) {
            ctx->worker_run = true;
            for (int i = 0; i < kIoSlots; i++) ctx->workers.emplace_back(worker_loop, ctx);
        }
        ctx->queue.push_back(std::move(r));
    }
    ctx->q_cv.notify_one();
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
        DeviceState d;
        d.id = id;
        d.sm_count = prop.multiProcessorCount;
        CU_TRY(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
        for (int w = 0; w < kSets; w++) {
            CU_TRY(cudaMalloc(&d.jobs_dev[w], lb2::kMaxJobs * sizeof(lb2::LayerJob)));
            CU_TRY(cudaMallocHost(&d.h_jobs[w], lb2::kMaxJobs * sizeof(lb2::LayerJob)));
            for (int i = 0; i < 8; i++) d.plan_key[w][i] = -1;
        }
        CU_TRY(cudaMalloc(&d.item_counter, 3 * sizeof(uint32_t)));   // [0] all items, [1 + net] resident-weights mode
        CU_TRY(cudaMemset(d.item_counter, 0, 3 * sizeof(uint32_t)));
        CU_TRY(cudaEventCreateWithFlags(&d.ev_user, cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&d.ev_comp, cudaEventDisableTiming));
        CU_TRY(lb2::trunk_kernel_setup());
        if (ctx->backend.empty()) ctx->backend = std::string("B200 tcgen05: ") + prop.name;
        ctx->dev.push_back(d);
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
    for (auto& d : ctx->dev) {
        cudaSetDevice(d.id);
        cudaStreamSynchronize(d.stream);
        for (int k = 0; k < 2; k++) {
            NetDev& nd = d.net[k];
            for (auto& t : nd.trunk)
                for (int sp = 0; sp < lb2::kMaxSplit; sp++) { cudaFree(t.wpk[sp]); cudaFree(t.wpk2[sp]); cudaFree(t.wpk_x[sp]); cudaFree(t.wpk2_x[sp]); cudaFree(t.bias[sp]); }
            for (int sp = 0; sp < lb2::kMaxSplit; sp++) cudaFree(nd.head_wt[sp]);
            cudaFree(nd.head_b); cudaFree(nd.ip1_wt); cudaFree(nd.ip1_b);
            cudaFree(nd.ip2_w); cudaFree(nd.ip2_b);
            free_workspace(&nd);
        }
        for (auto& sl : d.slots) free_slot(&sl);
        if (d.ev_user) cudaEventDestroy(d.ev_user);
        if (d.ev_comp) cudaEventDestroy(d.ev_comp);
        cudaFree(d.rot); cudaFree(d.item_counter);
        for (int w = 0; w < kSets; w++) { cudaFree(d.jobs_dev[w]); cudaFreeHost(d.h_jobs[w]); }
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
    }
    for (auto& d : net->ctx->dev) {
        CU_TRY(cudaSetDevice(d.id));
        int rc = upload_net(net, &d.net[net->kind]);
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
    // feature planes on the host's cores: ~0.1 ms per position each
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
    std::lock_guard<std::mutex> lk(ctx->eval_mu);
    DeviceState* d = &ctx->dev[dev_index];
    CU_TRY(cudaSetDevice(d->id));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : d->stream;
    const int chunk = (int)ctx->max_batch;
    for (int k = 0; k < 2; k++)
        if (need[k] && d->net[k].cap < std::min(n, chunk)) {
            CU_TRY(cudaDeviceSynchronize());   // nothing may be using the workspace while it grows
            if ((rc = ensure_workspace(&d->net[k], k, std::min(n, chunk)))) return rc;
        }
    if (st != d->stream) {   // the workspace is shared with host-buffer calls running on the compute stream
        CU_TRY(cudaEventRecord(d->ev_comp, d->stream));
        CU_TRY(cudaStreamWaitEvent(st, d->ev_comp, 0));
    }
    for (int lo = 0; lo < n; lo += chunk) {
        const int c = std::min(chunk, n - lo);
        rc = eval_on_device(ctx, d, d_pol ? d_pol + (size_t)lo * lb2::kPoints : nullptr,
                            d_val ? d_val + (size_t)lo * lb2::kPoints : nullptr, d_rot + lo, c, temp,
                            d_probs ? d_probs + (size_t)lo * lb2::kPoints : nullptr, d_win ? d_win + lo : nullptr, st, 0);
        if (rc) return rc;
    }
    if (st != d->stream) {
        CU_TRY(cudaEventRecord(d->ev_user, st));
        d->user_pending = true;
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
    ctx->q_idle.wait(lk, [&] { return ctx->queue.empty() && ctx->workers_busy == 0; });
    return LB2_OK;
}

const char* lb2_backend_name(lb2_ctx* ctx) { return ctx ? ctx->backend.c_str() : ""; }
int lb2_device_count(lb2_ctx* ctx) { return ctx ? (int)ctx->dev.size() : 0; }

int lb2_set_option(lb2_ctx* ctx, const char* name, long value) {
    if (!ctx || !name) return fail(LB2_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(ctx->eval_mu);
    if (!strcmp(name, "trunk_mode")) {
        if (value != 0 && value != 1) return fail(LB2_ERR_INVALID, "trunk_mode must be 0 or 1");
        ctx->trunk_mode = value;
    } else if (!strcmp(name, "trace")) {
        // debug: per-item timeline of the trunk kernel on device 0 (read back with lb2_debug_read_trace)
        DeviceState& d = ctx->dev[0];
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
    } else if (!strcmp(name, "cta_pair")) {
        ctx->cta_pair = value ? 1 : 0;
    } else if (!strcmp(name, "dynamic_items")) {
        ctx->dynamic_items = value ? 1 : 0;
    } else if (!strcmp(name, "overlap_io")) {
        ctx->overlap_io = value ? 1 : 0;
    } else if (!strcmp(name, "resident_weights")) {
        if (value < 0 || value > 2) return fail(LB2_ERR_INVALID, "resident_weights must be 0, 1 or 2");
        ctx->resident_weights = value;
    } else if (!strcmp(name, "precise")) {
        ctx->precise = value ? 1 : 0;
    } else if (!strcmp(name, "policy_clusters")) {
        ctx->policy_clusters = value;
    } else if (!strcmp(name, "profile_trunk")) {
        ctx->profile_trunk = value;
    } else if (!strcmp(name, "max_batch")) {
        if (value < 1 || value > 65536) return fail(LB2_ERR_INVALID, "max_batch out of range");
        ctx->max_batch = value;
    } else {
        return fail(LB2_ERR_INVALID, "unknown option %s", name);
    }
    return LB2_OK;
}

long lb2_get_option(lb2_ctx* ctx, const char* name) {
    if (!ctx || !name) return -1;
    if (!strcmp(name, "trunk_mode")) return ctx->trunk_mode;
    if (!strcmp(name, "max_batch")) return ctx->max_batch;
    if (!strcmp(name, "cta_pair")) return ctx->cta_pair;
    if (!strcmp(name, "dynamic_items")) return ctx->dynamic_items;
    if (!strcmp(name, "overlap_io")) return ctx->overlap_io;
    if (!strcmp(name, "resident_weights")) return ctx->resident_weights;
    if (!strcmp(name, "precise")) return ctx->precise;
    if (!strcmp(name, "policy_clusters")) return ctx->policy_clusters;
    if (!strcmp(name, "sm_count")) return ctx->dev.empty() ? 0 : ctx->dev[0].sm_count;
    if (!strcmp(name, "stat_positions")) return ctx->stat_positions.load();
    if (!strcmp(name, "stat_batches")) return ctx->stat_batches.load();
    if (!strcmp(name, "stat_requests")) return ctx->stat_requests.load();
    if (!strncmp(name, "seg", 3) && name[3] >= '0' && name[3] <= '3') {
        // mean ns of segment k over the evals recorded with profile_trunk == 2:
        // seg0 expand, seg1 trunk, seg2 heads; "seg3" only clears the record
        std::lock_guard<std::mutex> lk(ctx->eval_mu);
        DeviceState& d = ctx->dev[0];
        cudaSetDevice(d.id);
        cudaDeviceSynchronize();
        const int k = name[3] - '0';
        double ms_total = 0; long cnt = 0;
        for (size_t i = 0; k < 3 && i + 3 < d.seg_events.size(); i += 4) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, d.seg_events[i + k], d.seg_events[i + k + 1]) == cudaSuccess) { ms_total += ms; cnt++; }
        }
        if (k == 3) { for (auto e : d.seg_events) cudaEventDestroy(e); d.seg_events.clear(); }
        return cnt ? (long)(ms_total * 1e6 / cnt) : 0;
    }
    if (!strcmp(name, "trunk_ns") || !strcmp(name, "trunk_launches_timed")) {
        // device time spent in trunk launches since the last query (profile_trunk = 1); resets
        std::lock_guard<std::mutex> lk(ctx->eval_mu);
        double ms_total = 0;
        long pairs = 0;
        for (auto& d : ctx->dev) {
            cudaSetDevice(d.id);
            for (size_t i = 0; i + 1 < d.prof_events.size(); i += 2) {
                float ms = 0;
                cudaEventSynchronize(d.prof_events[i + 1]);
                if (cudaEventElapsedTime(&ms, d.prof_events[i], d.prof_events[i + 1]) == cudaSuccess) ms_total += ms;
                cudaEventDestroy(d.prof_events[i]);
                cudaEventDestroy(d.prof_events[i + 1]);
                pairs++;
            }
            d.prof_events.clear();
        }
        return !strcmp(name, "trunk_ns") ? (long)(ms_total * 1e6) : pairs;
    }
    return -1;
}

long lb2_launch_count(lb2_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }

int lb2_debug_read_trace(lb2_ctx* ctx, unsigned long long* out, long max_entries) {
    if (!ctx || !out) return fail(LB2_ERR_INVALID, "null argument");
    DeviceState& d = ctx->dev[0];
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
    std::lock_guard<std::mutex> lk(ctx->eval_mu);
    DeviceState* d = &ctx->dev[0];
    CU_TRY(cudaSetDevice(d->id));
    NetDev& nd = d->net[kind];
    if (n_layers < 1 || n_layers > (int)nd.trunk.size()) return fail(LB2_ERR_INVALID, "n_layers out of range");
    if ((rc = ensure_device_staging(d, n))) return rc;
    if ((rc = ensure_workspace(&nd, kind, n))) return rc;
    CU_TRY(cudaMemcpyAsync(d->rot, rotation, n, cudaMemcpyHostToDevice, d->stream));
    CU_TRY(cudaMemcpyAsync(nd.planes, planes, (size_t)n * lb2::kPoints * sizeof(uint32_t), cudaMemcpyHostToDevice, d->stream));
    lb2::ExpandArgs ea;
    memset(&ea, 0, sizeof ea);
    ea.rotation = d->rot; ea.n = n; ea.n_nets = 1;
    ea.planes[0] = nd.planes; ea.x0[0] = nd.x0[0]; ea.chunk_rows[0] = nd.rows5;
    CU_TRY(lb2::launch_expand(ea, d->stream));
    ctx->launches++;
    int limit[2] = {0, 0};
    limit[kind] = n_layers;
    JobPlan pl;
    if ((rc = run_trunk(ctx, d, need, n, limit, d->stream, &pl, 0))) return rc;
    const int c_out = nd.trunk[n_layers - 1].c_out;
    std::vector<__half> host((size_t)(c_out / 8) * nd.rows3 * 8);
    CU_TRY(cudaMemcpyAsync(host.data(), pl.last_act[kind], host.size() * sizeof(__half), cudaMemcpyDeviceToHost, d->stream));
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
