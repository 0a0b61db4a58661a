// lb2_kernels.cuh — device-side data model shared by the kernels and the host API.
//
// HBM layout of activations ("row space"):
//   A position's 19x19 feature map is stored as S x S rows (S = 20 for inputs of 3x3 convs,
//   S = 21 for the input of the first 5x5 conv): row = pos*S*S + y*S + x. Columns x >= 19 and
//   rows y >= 19 are zero padding SHARED between neighbours (the right pad of one board row is
//   the left pad of the next; the bottom pad of one position is the top pad of the next), so a
//   conv tap (dy, dx) of output row r reads input row r + dy*S + dx with no bounds logic.
//   Channels are split into chunks of 8 (16 bytes of fp16); a buffer is
//   [C/8 chunks][rows][8] fp16, i.e. each chunk plane is a dense array of 16-byte rows. This is
//   exactly the tcgen05 K-major no-swizzle core-matrix layout, so a [rows x 16 channels] slab
//   lands in shared memory with one TMA box and ANY row shift of it is a valid MMA operand.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace lb2 {

constexpr int kPoints = 361;
constexpr int kBoard = 19;
constexpr int kTileRows = 256;  // output rows per work item (two M=128 MMA halves)
constexpr int kStagesSingle = 3;  // smem ring depth (A slab + B block per stage), one CTA per item
constexpr int kStagesPair = 6;    // CTA pairs stage only half of B, so the ring is deeper
constexpr int kMaxHalo = 48;    // 5x5 conv in S=21 space needs 2*21+2 = 44 -> 48 (multiple of 8)
constexpr int kASlabBytes = (kTileRows + 2 * kMaxHalo) * 32;  // rows x 16 ch x fp16
constexpr int kBBlockBytes = 9 * 128 * 32;                    // up to 9 taps x N<=128 x 16 ch x fp16
constexpr int kStageBytesSingle = kASlabBytes + kBBlockBytes;      // 48,128
constexpr int kStageBytesPair = kASlabBytes + kBBlockBytes / 2;    // 29,696
constexpr int kMaxStages = 6;
#ifndef LB2_EPI_WARPS
#define LB2_EPI_WARPS 8
#endif
constexpr int kEpilogueWarps = LB2_EPI_WARPS;  // 8 or 16: TMEM lane quadrant (w % 4) x column part (w / 4)
constexpr int kColParts = kEpilogueWarps / 4;     // parts the output channels of a tile are split into
// warp0 TMA producer, warp1 MMA issuer, warps2-9 epilogue, warp10 tile publisher, warp11 dependency scout
constexpr int kTrunkThreads = 64 + 32 * kEpilogueWarps + 64;
constexpr int kMaxLaunchJobs = 44;  // jobs whose biases are kept resident in smem (12 + 11 layers, all but the first policy layer and the last ones cut in two)
constexpr int kMaxSplit = 2;        // a layer wider than 128 channels runs as this many column-split jobs
constexpr int kHeadSlots = 2 * kMaxSplit;  // fused-head weight sets resident in smem: [net][split]
constexpr int kPubDepth = 4;        // tiles that may be waiting for publication
// stages | mbarriers, TMEM slot, progress counters (256 B)
constexpr int kRingBytesSingle = kStagesSingle * kStageBytesSingle;   // 144,384
constexpr int kRingBytesPair = kStagesPair * kStageBytesPair;         // 178,176
constexpr int kTrunkRingBytes = kRingBytesSingle > kRingBytesPair ? kRingBytesSingle : kRingBytesPair;
// ring | mbarriers etc. | bias [jobs][128] f32 | fused-head weights [kHeadSlots][9][128] f32 | job table
constexpr int kCtrlBytes = 384;  // mbarriers, TMEM slot, progress counters, claim ring
constexpr int kTrunkSmemBytes = kTrunkRingBytes + kCtrlBytes + kMaxLaunchJobs * 128 * 4 + kHeadSlots * 9 * 128 * 4 + kMaxLaunchJobs * 192;
static_assert(kTrunkSmemBytes <= 227 * 1024, "trunk kernel shared memory");
// Resident-weights mode (CTA pairs only): a CTA keeps ITS half of the whole layer's packed weights in
// shared memory and re-uses it for every item of that layer it processes; the pipeline stages then
// carry activation slabs only. Needs c_in <= 128 and c_out <= 128 (9 x 128 x 64 x 2 B = 147,456 B).
constexpr int kResWeightBytes = 9 * 128 * 64 * 2;
constexpr int kStagesRes = 5;       // ring slots for jobs with the full halo (the 5x5 first layer): kASlabBytes each
constexpr int kResHalo3 = 24;       // halo of the 3x3 layers in S = 20 space (20 + 1 -> 24)
#ifndef LB2_RES_STAGES3
#define LB2_RES_STAGES3 6
#endif
constexpr int kStagesRes3 = LB2_RES_STAGES3;   // ... and for the 3x3 layers: (256 + 2 * 24) rows x 32 B = 9,728 B each
constexpr int kResRingBytes = (kStagesRes3 * (kTileRows + 2 * kResHalo3) * 32 > kStagesRes * kASlabBytes) ? kStagesRes3 * (kTileRows + 2 * kResHalo3) * 32
                                                                                                          : kStagesRes * kASlabBytes;
static_assert(kStagesRes3 <= kMaxStages && ((kTileRows + 2 * kResHalo3) * 32) % 128 == 0, "resident ring geometry");
constexpr int kResJobs = 24;        // jobs per launch in this mode (no column splits)
constexpr int kResHeadSlots = 2;    // one fused-head weight set per net
constexpr int kTrunkSmemBytesRes = kResWeightBytes + kResRingBytes + kCtrlBytes + kResJobs * 128 * 4 +
                                   kResHeadSlots * 9 * 128 * 4 + kResJobs * 192;
static_assert(kTrunkSmemBytesRes <= 227 * 1024, "resident-weights trunk shared memory");
constexpr int kMaxLayers = 16;
constexpr int kMaxTensorMaps = 6;
constexpr int kMaxJobs = 44;
constexpr int kMaxRounds = 104;  // a round = the jobs of equal depth (all nets, all column splits), their items interleaved; or, in
                                // net-major order, one position group of one layer ((12 + 11 layers) x up to 4 groups)
constexpr int kMaxRoundJobs = 2 * kMaxSplit;
constexpr int kTraceItems = 96;
constexpr int kTraceEvents = 16;
constexpr int kSchedEpoch = 3, kSchedWords = 4;
constexpr int kOutPlain = 0, kOutLo16 = 1, kOutFp8 = 2;
#ifndef LB2_LITE_SEPARATE_ACC
#define LB2_LITE_SEPARATE_ACC 0
#endif
constexpr bool kLiteSeparateAcc = LB2_LITE_SEPARATE_ACC != 0;
constexpr int kCorrCols = 64;   // lite mode: TMEM columns between an accumulator and the one of its e4m3 correction terms (c_out <= 64)
constexpr float kLoScale = 4096.0f;   // lite mode: the activation residual is stored as e4m3((a - fp16(a)) * 2^12)

// One trunk layer of one net over the whole batch.
struct LayerJob {
    int32_t n_items;         // work items: 256-row tiles, or pairs of tiles in CTA-pair mode
    int32_t item_base;       // index of its first item in the launch-wide item order
    int32_t S;               // row-space stride (20 or 21)
    int32_t ksize;           // 3 or 5
    int32_t halo;            // halo rows loaded each side of a tile (multiple of 8)
    int32_t n_slabs;         // c_in / 16
    int32_t n_out;           // output channels of this job = MMA N (multiple of 32, <= 128); a wider layer is split
    int32_t tmap;            // index of the input buffer's tensor map
    int32_t remap;           // 1: outputs are re-addressed from S=21 space into S=20 space
    int32_t dep_job;         // first job producing this job's input in the same launch, or -1
    int32_t dep_n_split;     // ... and how many consecutive (column-split) jobs produce it
    int32_t dep_remap;       // their remap flag
    int32_t dep_n_items;     // that job's tile count
    int32_t n_pos;           // positions in the batch
    int32_t out_chunk_rows;  // rows per chunk plane of the output buffer
    int32_t head_taps;       // 9: the net's final 3x3 conv to 1 channel is fused into this job's epilogue
    int32_t net;             // 0 = policy, 1 = value
    int32_t head_slot;       // which resident fused-head weight set (net * kMaxSplit + split)
    int32_t zparts;          // fused head: zbuf parts written before this job's (split * kColParts)
    int32_t layer;           // layer index within the net
    int32_t n_real_slabs;    // c_in / 16. Equal to n_slabs except in the split-operand modes, where the K loop runs over
                             // virtual slabs, term after term (see lb2_api.cu, pack_trunk_weights):
                             //   precise: [hi x Wh] [hi x Wl] [lo x Wh], all kind::f16
                             //   lite:    [hi x Wh] kind::f16, then [a8 | lo8] x [Wl8 ; W8] kind::f8f6f4 (e4m3, K = 32)
    int32_t lo_chunks;       // chunk planes of the OUTPUT buffer in front of the extra planes this job's epilogue stores
    int16_t term_base[3];    // virtual slab v = term * n_real_slabs + s reads input chunk planes term_base[term] + 2 s
    int16_t n_f16_slabs;     // virtual slabs [0, n_f16_slabs) are kind::f16 MMAs, the rest kind::f8f6f4
    int32_t group_tiles;     // net-major order runs the batch in position groups (all layers of group 0, then of group 1, ...) so that
    int32_t dep_group_tiles; // the live activations fit in L2: 256-row tiles per group of this job / of the producing job (0 = one group).
                             // A group's last tile does not wait for the next group's first one (it would read only padding rows of it)
    int32_t out_mode;        // what the epilogue stores besides the fp16 activations (for the consuming layer):
                             //   kOutPlain nothing, kOutLo16 the fp16 residual planes (precise consumer),
                             //   kOutFp8 per 16 channels one e4m3 plane of the activations and one of the residual x 2^12
    float acc_scale;         // the accumulator is acc_scale^-1 x the true sum (weights are packed scaled); 1 in plain mode
    const __half* wpk;       // packed weights: per (slab, tap group): [tap][2 chunks][n_out][8]
    const __half* wpk2;      // CTA-pair packing: per (slab, tap group): [rank][tap][2 chunks][n_out/2][8]
    const float* bias;       // [n_out]
    __half* out;             // output activation buffer
    uint32_t* flags;         // [n_items] completion flags (value = launch epoch)
    const __half* in_base;   // input buffer, for discarding dead tiles: in_planes chunk planes of in_chunk_rows rows each are dropped
    int32_t in_chunk_rows;   // from L2 once every tile of this layer that reads them is complete (0 planes = never: the first
    int32_t in_planes;       // layer's x0, column-split layers, per-layer launches)
    const float* head_w;     // fused head weights [9 taps][n_out] fp32
    float* zbuf;             // fused head output [splits * kColParts channel parts][9 taps][out_chunk_rows] fp32
};

struct TrunkParams {
    CUtensorMap tmaps[kMaxTensorMaps];
    LayerJob jobs[kMaxJobs];       // the launch's job table travels in the kernel parameters (no device copy to keep in step)
    int32_t n_jobs;
    int32_t item_begin, item_end;  // launch-wide item index range handled by this launch
    int32_t n_rounds;
    int32_t round_base[kMaxRounds + 1];             // first item index of each round
    int16_t round_first[kMaxRounds], round_jobs[kMaxRounds];  // its jobs: round_jobs consecutive entries of the job table
    int16_t round_idx0[kMaxRounds];                 // single-job rounds: the round covers items [idx0, idx0 + its length) of the job
    // Per-device scheduler words, reset by the expand kernel at the start of every evaluation (so that a launch does not
    // depend on its history and can be replayed from a CUDA graph): [0] in-order claim counter over all items, [1 + net] the
    // per-net claim counters of the resident-weights mode, [kSchedEpoch] the flag epoch of the evaluation.
    uint32_t* sched;
    int32_t use_flags;  // 1: cross-CTA dataflow through flags (single persistent launch)
    int32_t dynamic;            // 1: clusters claim items from sched[0] (or sched[1 + net]) in order; 0: static round robin
    uint32_t claim_base;        // sched[0] when this launch starts: per-layer launches (trunk_mode 0) share the counter, every
                                // launch advancing it by items + clusters
    // resident-weights mode: the item list is net-major (all policy jobs, then all value jobs), each net has
    // its own claim counter; a cluster works on its preferred net until that runs dry, then helps the other
    int32_t net_item_begin[2], net_item_end[2];
    int32_t policy_clusters;    // clusters [0, policy_clusters) prefer net 0, the rest net 1
    unsigned long long* trace;  // optional timeline buffer [cta][kTraceItems][kTraceEvents] of %globaltimer ns
    int32_t debug_flags;  // timing experiments only (results wrong): bit1 = all tap offsets 0, bit2 = no epilogue math,
                          // bit3 = no B loads, bit4 = no A loads, bit5 = only the first M half of 3x3 layers, bit6 = epilogue does nothing
};

struct ExpandArgs {
    const uint32_t* planes[2];   // packed bit-planes of up to two nets
    __half* x0[2];               // their first-conv input buffers (S=21 row space)
    int32_t chunk_rows[2];
    const uint8_t* rotation;     // [n] symmetry per position; ignored when `ensemble`
    int32_t ensemble;            // 1: device position p is input position p / 8 under symmetry p % 8 (AVERAGE_ALL)
    int32_t n, n_nets;
    const uint8_t* pf;           // optional buffer to pull into L2
    size_t pf_bytes;
    uint32_t* sched;             // the device's scheduler words (TrunkParams::sched): new epoch, claim counters to zero
};

struct HeadArgs {
    // policy head (n_policy positions, 0 = skip)
    const float* p_zbuf; int32_t p_chunk_rows; const float* p_bias; const uint8_t* rotation; float temp; float* probs;
    int32_t n_policy; int32_t p_parts;   // p_parts: zbuf parts to sum (column splits x kColParts)
    // value head (n_value positions, 0 = skip)
    const float* v_zbuf; int32_t v_chunk_rows; const float* v_bias; const float* ip1_wt; const float* ip1_b; int32_t hidden;
    const float* ip2_w; const float* ip2_b; float* winrate; int32_t n_value; int32_t v_parts;
    int32_t ensemble;            // 1: position p was evaluated under symmetry p % 8 (rotation is unused)
    float* v_partial;            // value head workspace: per (group of 16 positions, slice of the 361 x H matrix) partial sums [16][256]
    uint32_t* v_count;           // ... and per group the number of slices done (zero between launches)
    unsigned long long* trace;   // debug timeline (option "trace"): block b < trace_ctas stamps %globaltimer at its start and end
    int32_t trace_ctas;          // into slot [b][kTraceItems - 2][0 / 8] (tools/trace_heads.py)
};

// AVERAGE_ALL on the device: mean over the 8 symmetries of each position, summed in the reference's
// order r = 0..7 then divided by 8 (Network.cpp:605-615, 643-654)
struct MeanArgs {
    const float* probs8; float* probs; int32_t n_policy;   // [8n][361] -> [n][361]
    const float* win8; float* win; int32_t n_value;        // [8n] -> [n]
};

// launchers (lb2_kernels.cu)
cudaError_t launch_expand(const ExpandArgs& a, cudaStream_t st);
cudaError_t launch_trunk(const TrunkParams& p, int grid, bool cooperative, bool pair, bool resident, int out_modes, cudaStream_t st);
cudaError_t launch_heads(const HeadArgs& a, cudaStream_t st);
size_t heads_partial_floats(int n_value);   // sizes of HeadArgs::v_partial / v_count for up to n_value positions
size_t heads_count_words(int n_value);
cudaError_t launch_ensemble_mean(const MeanArgs& a, cudaStream_t st);
cudaError_t trunk_kernel_setup();
const void* kernel_address(int which);   // 0 expand, 1 heads, 2 ensemble mean

}  // namespace lb2
static_assert(sizeof(lb2::TrunkParams) < 32000, "kernel parameter space");
static_assert(sizeof(lb2::LayerJob) <= 192 && sizeof(lb2::LayerJob) % 4 == 0, "job table slot size");
