/* leela_b200.h — C ABI of the B200 (sm_100a) evaluator for Leela's policy and value networks.
 *
 * This is the drop-in boundary for ONE hot path of gcp/Leela: the convolutional policy+value
 * network evaluation behind Network.cpp. It replaces the reference's two backends
 *   - OpenCL:  OpenCL.h:47-134 (OpenCL_Network::push_convolve / push_innerproduct / forward,
 *              OpenCL::initialize / join_outstanding_cb), kernels OpenCL.cpp:25-438
 *   - CPU:     Im2Col.h:8-50 + cblas_sgemm/sgemv templates, Network.cpp:344-447
 * and is what Network::get_scored_moves / get_value / async_scored_moves (Network.h:43-61)
 * call instead. Plain pointers and sizes only; no C++ or torch types cross this boundary.
 *
 * Conventions
 *   - Every function returns LB2_OK (0) or a negative lb2_status; nothing throws.
 *     lb2_last_error() gives the text of the calling thread's most recent failure.
 *   - Board is 19x19. A position's input is 32 binary feature planes packed one uint32 per
 *     board point: planes[i*361 + idx], idx = y*19 + x, bit c = plane c, in the order of
 *     Network::gather_features_policy (Network.cpp:886-917) or gather_features_value
 *     (Network.cpp:1050-1081). (std::bitset<361> x 32 -> uint32[361].)
 *   - rotation[i] in 0..7 is the symmetry of Network::rotate_nn_idx (Network.cpp:1348-1379)
 *     the position is evaluated under; outputs come back un-rotated (rev_rotate_nn_idx,
 *     Network.cpp:820-823), i.e. indexed by the ORIGINAL board idx.
 *   - Policy output is the temperature softmax over all 361 points (Network::softmax,
 *     Network.cpp:450-469). Filtering to EMPTY points, vertex mapping and losing-ladder
 *     pruning (Network.cpp:820-829, 656-667) need the board and stay with the caller.
 *   - Value output is (1 + tanh(x)) / 2 for the side to move (Network.cpp:736).
 *   - Host pointers are caller-owned and may be pageable; pinned staging is internal.
 *   - There is no CPU fallback: without a B200 and the compiled kernels, lb2_init fails.
 */
#ifndef LEELA_B200_H
#define LEELA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LB2_BOARD_POINTS 361
#define LB2_INPUT_PLANES 32

typedef enum lb2_status {
    LB2_OK = 0,
    LB2_ERR_INVALID = -1,     /* bad argument: n < 0, rotation > 7, null pointer, bad geometry */
    LB2_ERR_CUDA = -2,        /* a CUDA call failed (message in lb2_last_error) */
    LB2_ERR_STATE = -3,       /* net not finalized / already finalized / wrong net kind */
    LB2_ERR_NOMEM = -4,
    LB2_ERR_UNSUPPORTED = -5  /* layer stack this evaluator has no kernel for */
} lb2_status;

typedef enum lb2_net_kind { LB2_POLICY = 0, LB2_VALUE = 1 } lb2_net_kind;

typedef struct lb2_ctx lb2_ctx;
typedef struct lb2_net lb2_net;

/* Replaces OpenCL::initialize (OpenCL.cpp:764-908). device_ids == NULL or n_devices == 0
 * selects CUDA device 0. With several devices, weights are replicated on each; a call of up to
 * max_batch positions runs as ONE batch on whichever device has a free I/O slot, larger calls
 * are cut into max_batch chunks that are dealt to the devices as they become free (no collective,
 * and never slivers of one batch on every device). */
int lb2_init(const int* device_ids, int n_devices, lb2_ctx** ctx_out);
void lb2_destroy(lb2_ctx* ctx);

/* Replaces the globals opencl_policy_net / opencl_value_net (OpenCL.h:137-138): one net of
 * each kind per context. The net is owned by the context. */
int lb2_net_create(lb2_ctx* ctx, int kind, lb2_net** net_out);

/* Replaces OpenCL_Network::push_convolve (OpenCL.h:66-87): w is OIHW fp32
 * [c_out][c_in][k][k] (Network.cpp:363), bias [c_out]. Bias + ELU follow every conv
 * (Network.cpp:382-392). Layers run in push order (Network.cpp:206-233). */
int lb2_net_push_conv(lb2_net* net, int k, int c_in, int c_out, const float* w_oihw, const float* bias);

/* Replaces OpenCL_Network::push_innerproduct (OpenCL.h:89-94): w row-major [n_out][n_in];
 * ELU follows iff n_out > 1 (Network.cpp:411-421). */
int lb2_net_push_ip(lb2_net* net, int n_in, int n_out, const float* w, const float* bias);

/* Repacks the weights for the tensor-core kernels and replicates them to every device.
 * Supported stacks: a 5x5 conv from 32 planes, then 3x3 convs (c_out a multiple of 32 up to 128,
 * or a multiple of 64 up to 256 — run as two column splits), then a 3x3 conv to 1 channel; the value
 * net adds innerproduct 361 -> H (H <= 256) and H -> 1. That covers NN128 (Network.cpp:82-107),
 * the 192-wide OpenCL-build policy net (Network.cpp:55-80) and NNValue (Network.cpp:110-137). */
int lb2_net_finalize(lb2_net* net);

/* Replace OpenCL_Network::forward (OpenCL.cpp:490-569) for a BATCH of n positions (the
 * reference evaluates one). Blocking. probs_out: [n][361] fp32; winrate_out: [n] fp32. */
int lb2_eval_policy(lb2_ctx* ctx, const uint32_t* planes, const uint8_t* rotation, int n,
                    float softmax_temp, float* probs_out);
int lb2_eval_value(lb2_ctx* ctx, const uint32_t* planes, const uint8_t* rotation, int n,
                   float* winrate_out);
/* Both nets for the same n positions (policy and value planes are different 32-plane sets),
 * run concurrently on the device. The benchmark unit. */
int lb2_eval_both(lb2_ctx* ctx, const uint32_t* policy_planes, const uint32_t* value_planes,
                  const uint8_t* rotation, int n, float softmax_temp,
                  float* probs_out, float* winrate_out);

/* Ensemble AVERAGE_ALL (Network.cpp:605-615 value, 643-654 policy) for n positions: each position
 * is expanded on the device under all 8 symmetries from ONE copy of its planes, the 8 un-rotated
 * results are summed in the reference's order (r = 0..7) and divided by 8 on the device — the
 * caller uploads 1/8 and downloads 1/8 of what an explicit 8-entry batch would. Either output (with
 * its input) may be NULL. Blocking. probs_out: [n][361]; winrate_out: [n]. */
int lb2_eval_ensemble(lb2_ctx* ctx, const uint32_t* policy_planes, const uint32_t* value_planes, int n,
                      float softmax_temp, float* probs_out, float* winrate_out);

/* Same as lb2_eval_both but every pointer is DEVICE memory on context device `dev_index`,
 * work is enqueued on `cuda_stream` (a cudaStream_t, NULL = the context's own stream) and the
 * call returns without synchronising. Used to time the kernels with inputs resident in HBM.
 * Either output pointer may be NULL to skip that net. */
int lb2_eval_both_device(lb2_ctx* ctx, int dev_index, const uint32_t* d_policy_planes,
                         const uint32_t* d_value_planes, const uint8_t* d_rotation, int n,
                         float softmax_temp, float* d_probs_out, float* d_winrate_out,
                         void* cuda_stream);

/* Asynchronous submission from many search threads; replaces forward(cb) +
 * thread_can_issue / join_outstanding_cb (OpenCL.cpp:446-454, 560-577). A request's planes are
 * copied straight into the pinned buffer of the batch that is filling up (so the input buffers are
 * free again when the call returns); two dispatcher threads per device take whatever has accumulated
 * as soon as they are free — the batch size follows the load — and run it as one device batch;
 * cb(user, status) runs on such a thread once the caller's output buffer is filled. When status is
 * not LB2_OK, lb2_last_error() inside the callback (or lb2_queue_error later, from any thread) gives
 * the text. A request of more than max_batch positions is evaluated at once on the calling thread
 * (its callback runs before lb2_submit_* returns). */
typedef void (*lb2_callback)(void* user, int status);
int lb2_submit_policy(lb2_ctx* ctx, const uint32_t* planes, const uint8_t* rotation, int n,
                      float softmax_temp, float* probs_out, lb2_callback cb, void* user);
int lb2_submit_value(lb2_ctx* ctx, const uint32_t* planes, const uint8_t* rotation, int n,
                     float* winrate_out, lb2_callback cb, void* user);
int lb2_drain(lb2_ctx* ctx);
/* Text of the most recent asynchronous failure (empty if none), copied into buf. */
int lb2_queue_error(lb2_ctx* ctx, char* buf, int len);

/* Optional: page-lock a caller buffer (cudaHostRegister) so that lb2_eval_* copy to / from it directly
 * instead of through the library's pinned staging; a buffer that is already page-locked (cudaMallocHost,
 * another library) is only remembered. Unregistered pointers are still recognised as pinned or pageable —
 * one cudaPointerGetAttributes the first time an address is seen — so this is about skipping even that. */
int lb2_register_host_buffer(lb2_ctx* ctx, void* ptr, size_t bytes);
int lb2_unregister_host_buffer(lb2_ctx* ctx, void* ptr);

/* Feature planes from a RAW position — replaces Network::gather_features_policy / _value
 * (Network.cpp:883-1201) together with the FastBoard queries they make (liberties, liberties after a
 * move, the ladder readers: FastBoard.cpp:2482-2564, 2647-2837), bit for bit, with a board of this
 * library's own. Pure host code (no device needed); thread-safe.
 *   stones[361]     idx = y*19 + x: 0 empty, 1 black, 2 white
 *   white_to_move   side to move
 *   ko_point        idx of the point forbidden by ko (FastState::get_komove), or -1
 *   last_move       idx of the last move, -1 if none or a pass; prev_move likewise (only used when
 *                   last_move >= 0, as in Network.cpp:1026-1036)
 *   komi            white's komi (plane "has_komi" is set for white stones when |komi| > 0.75)
 * policy_planes / value_planes: [361] packed as lb2_eval_* take them; either may be NULL. */
int lb2_planes_from_position(const uint8_t* stones, int white_to_move, int ko_point, int last_move, int prev_move,
                             float komi, uint32_t* policy_planes, uint32_t* value_planes);

/* Raw positions in, results out: lb2_planes_from_position on the host's cores (the positions are
 * spread over worker threads), then lb2_eval_both. For callers that have no Go board code at all.
 * rotation may be NULL (symmetry 0 for every position); either output may be NULL. */
typedef struct lb2_position {
    uint8_t stones[LB2_BOARD_POINTS];   /* idx = y*19 + x: 0 empty, 1 black, 2 white */
    uint8_t white_to_move;
    int16_t ko_point, last_move, prev_move;   /* idx, or -1 */
    float komi;
} lb2_position;
int lb2_eval_positions(lb2_ctx* ctx, const lb2_position* positions, const uint8_t* rotation, int n,
                       float softmax_temp, float* probs_out, float* winrate_out);

/* Replaces OpenCL::get_device_name / Network::get_backend (Network.cpp:1535-1553). */
const char* lb2_backend_name(lb2_ctx* ctx);
const char* lb2_last_error(void);
int lb2_device_count(lb2_ctx* ctx);

/* Tuning / introspection.
 *   "trunk_mode": 0 = one launch per layer, 1 = single persistent dataflow launch (default)
 *   "cta_pair":   1 = tensor-core work issued for CTA pairs (tcgen05 cta_group::2, default), 0 = per CTA
 *   "dynamic_items": 1 = clusters claim work items from a global in-order counter (default), 0 = round robin
 *   "resident_weights": each CTA keeps its half of a layer's packed weights in shared memory for all the layer's items
 *                 it processes, the clusters are split between the two nets (by estimated work; each helps the other
 *                 out when its own net runs dry) and the pipeline stages carry activation slabs only. 0 = off, 1 = whenever
 *                 every layer fits (no column splits; policy layers only in fp16 precision), 2 = only for launches that run
 *                 both nets (default: 4 % faster there, no gain for one net alone). Bit-identical to the streaming form.
 *   "policy_precision", "value_precision": arithmetic of a net's conv stack (accumulation, epilogue and heads are always fp32):
 *                 0 = fp16 operands, one tensor-core term per layer. Measured against the reference's fp32 OpenBLAS path over
 *                     the 1024-position correctness set: policy max |dp| 5.3e-3, value max 3.3e-3.
 *                 1 = "lite" split operands: the fp16 term plus ONE e4m3 (kind::f8f6f4, K = 32) correction term per 16
 *                     channels, [e4m3(a) | e4m3((a - fp16(a)) 2^12)] x [e4m3(w - fp16(w)) ; e4m3(w)], scaled by powers of two
 *                     into e4m3's range and accumulated in the same fp32 accumulator: twice the tensor work, value max 1.3e-4.
 *                 2 = "full" split operands: fp16 hi + fp16 lo for activations and weights, three fp16 terms (hi*Wh + hi*Wl +
 *                     lo*Wh): three times the tensor work, policy max 8.9e-5, value max 2.8e-5.
 *                 Default: policy 0, value 1 (north_star: value within 1e-3; the policy tolerance is stated from measurement).
 *                 Lite and full cannot be mixed between the two nets. May be switched between calls.
 *   "precise":    shorthand: 1 = both nets full (2, 2); 0 = the default (0, 1).
 *   "policy_clusters": resident mode: clusters that start on the policy net (-1 = split by estimated work)
 *   "small_batch": device passes of up to this many positions (default 48, 0 = never) run every layer of up to 128 channels as
 *                 two column-split jobs, so that two clusters share an item's MMAs and epilogue: a small pass is bound by the
 *                 latency of its chained layers, not by throughput (batch 1: 154 -> 131 us, batch 32: 170 -> 159 us; above ~56
 *                 positions whole layers are faster). Bit-identical.
 *   "group_positions": net-major launches (resident_weights) run the batch in groups of this many positions — all layers of
 *                 the first group, then of the next — so that the live activations fit in L2. A multiple of 128, 0 = off
 *                 (default: at batch 256 groups of 128 cut the HBM write-back from 389 to 94 MB per launch but cost 5 % time)
 *   "use_graphs": 1 = from the second use of a batch shape on, its kernels (expand, trunk, heads) are launched as one
 *                 cached CUDA graph (default); 0 = always as separate launches
 *   "spin_wait":  1 = a blocking call polls for its results, yielding the core between polls (default); 0 = it sleeps
 *                 on a blocking-sync event (frees the core, adds wake-up latency)
 *   "queue_linger": 1 = a dispatcher of the submit queue that holds a free I/O slot lets requests accumulate while the device
 *                 is still computing the previous batch (kernels of one device run one after the other, and a small pass
 *                 costs the same ~130 us whatever its size), until that pass is done or a full batch is waiting (default);
 *                 0 = it carries off whatever has arrived at once
 *   "max_batch":  positions per device pass (larger calls are chunked), default 256
 *   "profile_trunk": 1 = bracket every trunk launch with CUDA events (nodes of the graph when graphs are on);
 *                 lb2_get_option("trunk_ns") then returns the device nanoseconds accumulated since the last query.
 *                 "profile_reserve" = N creates N event pairs per device ahead of time (none is then created inside a timed loop)
 *   read-only: "stat_positions", "stat_batches", "stat_requests" = positions, device batches and
 *                 requests that went through lb2_submit_* so far (mean batch = positions / batches); "graph_launches" */
int lb2_set_option(lb2_ctx* ctx, const char* name, long value);
long lb2_get_option(lb2_ctx* ctx, const char* name);
/* Number of kernel launches issued by this context so far (bench.py's gpu_launches). */
long lb2_launch_count(lb2_ctx* ctx);

/* Test hook: run the first `n_layers` convs of net `kind` on n positions and return the
 * activations of the last one as fp32 [n][c_out][361] in NETWORK orientation (i.e. still
 * rotated), after bias + ELU, as stored (fp16-rounded). Lets tests check each layer against
 * the oracle's convolve<>. */
int lb2_debug_trunk(lb2_ctx* ctx, int kind, const uint32_t* planes, const uint8_t* rotation,
                    int n, int n_layers, float* act_out);

/* Debug hook: after lb2_set_option(ctx, "trace", 1), every trunk launch records a per-work-item
 * timeline (global-timer nanoseconds, [cta][96 items][16 events]); this copies it out and clears
 * it. Returns the number of CTAs copied, or a negative status. */
int lb2_debug_read_trace(lb2_ctx* ctx, unsigned long long* out, long max_entries);

#ifdef __cplusplus
}
#endif
#endif /* LEELA_B200_H */
