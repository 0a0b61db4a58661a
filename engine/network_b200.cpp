// network_b200.cpp — the reference's `Network` class (Network.h:25-110) implemented on the B200
// evaluator's C ABI (include/leela_b200.h). Compiled INSTEAD of the reference's Network.cpp and
// OpenCL.cpp into the drop-in engine (engine/Makefile); every caller of the scoring API —
// UCTNode.cpp:108-118, 334; UCTSearch.cpp:811, 901; GTP.cpp:111, 489-508, 618, 789 — is the
// reference's own, unmodified code.
//
// What lives here (host side of the boundary):
//   * feature planes      gather_features_policy / _value (Network.cpp:883-1201) re-expressed as one
//                         table-driven pass that writes the packed form the C ABI takes (one
//                         uint32 per board point, bit c = plane c) using the reference's FastBoard
//                         queries (count_rliberties, after_liberties, ladder readers ...)
//   * ensembles           DIRECT / RANDOM_ROTATION / AVERAGE_ALL (Network.cpp:590-674); AVERAGE_ALL is
//                         lb2_eval_ensemble: expanded x8, evaluated and averaged on the device
//   * result mapping      EMPTY filter, vertex mapping, losing-ladder prune (Network.cpp:820-829, 656-667)
//   * async expansion     async_scored_moves + completion callback -> UCTNode::scoring_cb
//                         (Network.cpp:471-588), on lb2_submit_policy
//   * the `opencl` seam   thread_can_issue / join_outstanding_cb ... (OpenCL.h:113-134, OpenCL.cpp:440-577)
// What does NOT live here: plane expansion, rotation, the conv stacks, softmax(T), un-rotation,
// inner products, tanh — those run on the GPU behind the C ABI.
//
// Every blocking evaluation goes through lb2_submit_* + a per-call waiter rather than
// lb2_eval_*: requests from all search threads are then coalesced by the library's worker into
// one device batch (the reference evaluates batch 1 per thread).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cassert>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "config.h"
#include "Network.h"
#include "FastBoard.h"
#include "FastState.h"
#include "GTP.h"
#include "Random.h"
#include "ThreadPool.h"
#include "Timing.h"
#include "UCTNode.h"
#include "Utils.h"

#include "leela_b200.h"
#include "network_b200.h"

using namespace Utils;

Network* Network::s_Net = nullptr;
OpenCL opencl;

namespace {

lb2_ctx* g_ctx = nullptr;
std::string g_weights_path;
int g_max_outstanding = 2;   // per search thread, as OpenCL::thread_can_issue (OpenCL.cpp:446-454)
int g_planes_mode = 1;       // 1 (default): lb2_planes_from_position (own board, 2-12x faster); 0: planes through the reference's board queries;
                             // 2: both, and abort on the first difference (cross-check on the positions a real search visits)
std::atomic<long> g_planes_checked{0};
thread_local std::atomic<int> t_results_outstanding{0};

struct KnownAnswer { int index, vertex; float prob; };
std::vector<KnownAnswer> g_kat;   // known answers shipped with the weights file, see self_test()
float g_kat_value = -1.0f;

[[noreturn]] void die(const char* what) {
    // the OpenCL backend threw std::runtime_error / cl::Error here (OpenCL.cpp:665-669, 848-850)
    throw std::runtime_error(std::string("leela_b200: ") + what + ": " + lb2_last_error());
}

// ------------------------------------------------------------------------------------------
// weights file ("LB2WGT01", written by leela_b200/fileio.py:write_weights): per net the conv and
// inner-product arrays in the layout of the reference's extern arrays (Network.cpp:54-137):
// conv OIHW, bias, ip [n_out][n_in].
// ------------------------------------------------------------------------------------------
struct FileReader {
    FILE* f;
    explicit FileReader(const std::string& path) : f(fopen(path.c_str(), "rb")) {
        if (!f) throw std::runtime_error("leela_b200: cannot open weights file " + path);
    }
    ~FileReader() { fclose(f); }
    void read(void* dst, size_t bytes) {
        if (fread(dst, 1, bytes, f) != bytes) throw std::runtime_error("leela_b200: weights file truncated");
    }
    int32_t i32() { int32_t v; read(&v, 4); return v; }
    std::vector<float> floats(size_t n) { std::vector<float> v(n); read(v.data(), n * 4); return v; }
};

void load_weights(const std::string& path) {
    FileReader r(path);
    char magic[8];
    r.read(magic, 8);
    if (memcmp(magic, "LB2WGT01", 8)) throw std::runtime_error("leela_b200: " + path + " is not a weights file");
    const int n_nets = r.i32();
    for (int i = 0; i < n_nets; i++) {
        const int kind = r.i32(), n_convs = r.i32(), n_ips = r.i32();
        lb2_net* net = nullptr;
        if (lb2_net_create(g_ctx, kind, &net)) die("lb2_net_create");
        for (int l = 0; l < n_convs; l++) {
            const int k = r.i32(), c_in = r.i32(), c_out = r.i32();
            const std::vector<float> w = r.floats((size_t)k * k * c_in * c_out), b = r.floats(c_out);
            if (lb2_net_push_conv(net, k, c_in, c_out, w.data(), b.data())) die("lb2_net_push_conv");
        }
        for (int l = 0; l < n_ips; l++) {
            const int n_in = r.i32(), n_out = r.i32();
            const std::vector<float> w = r.floats((size_t)n_in * n_out), b = r.floats(n_out);
            if (lb2_net_push_ip(net, n_in, n_out, w.data(), b.data())) die("lb2_net_push_ip");
        }
        if (lb2_net_finalize(net)) die("lb2_net_finalize");
    }
    // optional known-answer trailer (leela_b200/fileio.py:write_weights)
    if (fread(magic, 1, 8, r.f) == 8 && !memcmp(magic, "LB2KAT01", 8)) {
        const int n = r.i32();
        for (int i = 0; i < n; i++) {
            KnownAnswer k;
            k.index = r.i32(); k.vertex = r.i32();
            r.read(&k.prob, 4);
            g_kat.push_back(k);
        }
        r.read(&g_kat_value, 4);
    }
}

// ------------------------------------------------------------------------------------------
// feature planes. Plane numbering of the two nets (Network.cpp:886-917 policy, :1050-1081 value):
// planes 0..2 are empty / side to move / opponent in both; the rest is described by this table.
// ------------------------------------------------------------------------------------------
struct PlaneLayout {
    int own_libs, opp_libs, libs_cap;     // real liberties of the string: planes base .. base+cap-1 (last = ">= cap")
    int own_after, opp_after;             // liberties after playing here, 1..5 and >= 6
    int ladder, ladder_win, ko, last_move, prev_move, komi, line3;
};
const PlaneLayout kPolicyPlanes = {3, 8, 5, 13, 19, 25, 26, 27, 28, 29, 30, 31};
const PlaneLayout kValuePlanes = {3, 9, 6, 15, 21, 27, 28, 31, -1, -1, 29, 30};

inline uint32_t plane(int p) { return 1u << p; }
inline uint32_t count_plane(int base, int count, int cap) { return count >= 1 ? plane(base + std::min(count, cap) - 1) : 0u; }

}  // namespace

namespace leela_b200 {

// The same planes from the raw position through the library's own board (lb2_planes_from_position).
static void pack_features_own_board(FastState* state, bool value_net, uint32_t* packed) {
    FastBoard& board = state->board;
    uint8_t stones[361];
    for (int idx = 0; idx < 361; idx++) {
        const FastBoard::square_t sq = board.get_square(board.get_vertex(idx % 19, idx / 19));
        stones[idx] = sq == FastBoard::BLACK ? 1 : (sq == FastBoard::WHITE ? 2 : 0);
    }
    auto index_of = [&](int vertex) {
        if (vertex <= 0) return -1;   // none or pass
        const std::pair<int, int> xy = board.get_xy(vertex);
        return xy.second * 19 + xy.first;
    };
    if (lb2_planes_from_position(stones, state->get_to_move() == FastBoard::WHITE, index_of(state->get_komove()),
                                 index_of(state->get_last_move()), index_of(state->get_prevlast_move()), state->get_komi(),
                                 value_net ? nullptr : packed, value_net ? packed : nullptr))
        die("lb2_planes_from_position");
}

static void pack_features_reference_board(FastState* state, bool value_net, uint32_t* packed);

// packed[idx], idx = y*19 + x. `ladder` (optional) receives the losing-ladder plane, which the
// callers need on the host to prune the result (Network.cpp:656-667).
void pack_features(FastState* state, bool value_net, uint32_t* packed, Network::BoardPlane* ladder) {
    if (g_planes_mode == 1) {
        pack_features_own_board(state, value_net, packed);
    } else {
        pack_features_reference_board(state, value_net, packed);
        if (g_planes_mode == 2) {
            uint32_t own[361];
            pack_features_own_board(state, value_net, own);
            if (memcmp(own, packed, sizeof own)) {
                fprintf(stderr, "leela_b200: PLANE MISMATCH between the own-board builder and the reference's board queries (%s net, move %d)\n",
                        value_net ? "value" : "policy", (int)state->get_movenum());
                abort();
            }
            g_planes_checked++;
        }
    }
    if (ladder) {
        const int bit = value_net ? kValuePlanes.ladder : kPolicyPlanes.ladder;
        ladder->reset();
        for (int idx = 0; idx < 361; idx++)
            if ((packed[idx] >> bit) & 1u) ladder->set(idx);
    }
}

static void pack_features_reference_board(FastState* state, bool value_net, uint32_t* packed) {
    const PlaneLayout& L = value_net ? kValuePlanes : kPolicyPlanes;
    FastBoard& board = state->board;
    const int tomove = state->get_to_move();
    const bool white_has_komi = std::fabs(state->get_komi()) > 0.75f;
    for (int idx = 0; idx < 361; idx++) {
        const int x = idx % 19, y = idx / 19;
        const int vtx = board.get_vertex(x, y);
        const FastBoard::square_t sq = board.get_square(vtx);
        uint32_t bits = (x == 2 || x == 16 || y == 2 || y == 16) ? plane(L.line3) : 0u;
        if (sq != FastBoard::EMPTY) {
            const bool own = (sq == tomove);
            bits |= own ? plane(1) : plane(2);
            if (sq == FastBoard::WHITE && white_has_komi) bits |= plane(L.komi);   // white gets extra points in scoring
            bits |= count_plane(own ? L.own_libs : L.opp_libs, board.count_rliberties(vtx), L.libs_cap);
        } else {
            bits |= plane(0);
            const std::pair<int, int> after = board.after_liberties(tomove, vtx);
            bits |= count_plane(L.own_after, after.first, 6) | count_plane(L.opp_after, after.second, 6);
            // escaping move of a string in atari that still runs into a ladder
            if (board.count_pliberties(vtx) == 2 && board.saving_size(tomove, vtx) > 0 &&
                board.check_losing_ladder(tomove, vtx))
                bits |= plane(L.ladder);
            if (board.check_winning_ladder(tomove, vtx)) bits |= plane(L.ladder_win);
        }
        packed[idx] = bits;
    }
    auto mark = [&](int vertex, int p) {
        if (p < 0 || vertex <= 0) return false;
        const std::pair<int, int> xy = board.get_xy(vertex);
        packed[xy.second * 19 + xy.first] |= plane(p);
        return true;
    };
    if (mark(state->get_last_move(), L.last_move)) mark(state->get_prevlast_move(), L.prev_move);
    mark(state->get_komove(), L.ko);
}

// The reference checks its OpenCL backend at start-up against two known outputs of ITS weights on the empty board
// (GTP::perform_self_test, GTP.cpp:105-125: entries 60 and 72 of get_scored_moves(DIRECT, 0)). The same check for whatever
// weights were loaded: the file carries the answers of the reference's CPU path for them (probabilities of a few
// Netresult entries with their vertices, and the winrate), compared at the tolerances the parity tests state.
bool self_test(GameState& state) {
    if (g_kat.empty()) {
        myprintf("B200 self-test: skipped (the weights file carries no known answers).\n");
        return true;
    }
    myprintf("B200 self-test: ");
    bool ok = true;
    const Network::Netresult vec = Network::get_Network()->get_scored_moves(&state, Network::Ensemble::DIRECT, 0);
    for (const KnownAnswer& k : g_kat) {
        ok = ok && k.index >= 0 && k.index < (int)vec.size();
        if (!ok) break;
        ok = ok && vec[k.index].second == k.vertex && std::fabs(vec[k.index].first - k.prob) < 6e-3f;
    }
    const float win = Network::get_Network()->get_value(&state, Network::Ensemble::DIRECT);   // DIRECT = symmetry 0
    ok = ok && std::fabs(win - g_kat_value) < 1e-3f;
    myprintf(ok ? "passed.\n" : "failed. The evaluator does not reproduce the reference's outputs for these weights.\n");
    return ok;
}

void set_planes_mode(int mode) { g_planes_mode = mode; }
long planes_checked() { return g_planes_checked.load(); }
void set_weights_path(const std::string& path) { g_weights_path = path; }
void set_max_outstanding(int n) { g_max_outstanding = std::max(1, n); }
lb2_ctx* context() { return g_ctx; }

}  // namespace leela_b200

namespace {

// probs[361] (un-rotated, all points) -> scored moves over EMPTY points in raster order, then
// losing ladders zeroed (Network.cpp:820-829, 656-667). No renormalisation.
Network::Netresult to_netresult(FastState& state, const float* probs, const Network::BoardPlane& ladder) {
    Network::Netresult result;
    for (int idx = 0; idx < 361; idx++) {
        const int vtx = state.board.get_vertex(idx % 19, idx / 19);
        if (state.board.get_square(vtx) == FastBoard::EMPTY)
            result.emplace_back(ladder[idx] ? 0.0f : probs[idx], vtx);
    }
    return result;
}

// one blocking evaluation = submit + wait, so that concurrent callers share a device batch
struct Waiter {
    std::mutex mu;
    std::condition_variable cv;
    bool done = false;
    int status = 0;
    std::string text;   // the failure's text: lb2_last_error() is per thread, and the callback runs on the dispatcher's
    static void signal(void* user, int status) {
        Waiter* w = static_cast<Waiter*>(user);
        std::lock_guard<std::mutex> lk(w->mu);
        w->status = status;
        if (status) w->text = lb2_last_error();
        w->done = true;
        w->cv.notify_one();
    }
    void wait(const char* what) {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return done; });
        if (status) {
            myprintf("leela_b200: %s failed: %s\n", what, text.c_str());
            throw std::runtime_error(std::string("leela_b200: ") + what + ": " + text);
        }
    }
};

// completion of an asynchronous policy expansion (the reference's CallbackData + forward_cb,
// Network.cpp:471-534)
struct AsyncExpansion {
    std::atomic<int>* nodecount;
    FastState state;
    UCTNode* node;
    Network::BoardPlane ladder;
    std::atomic<int>* thread_outstanding;
    uint32_t planes[361];
    uint8_t rotation;
    float probs[361];
};

void async_done(void* user, int status) {
    AsyncExpansion* x = static_cast<AsyncExpansion*>(user);
    x->thread_outstanding->fetch_sub(1, std::memory_order_release);   // the issuing thread may queue again
    if (status == LB2_OK) {
        Network::Netresult result = to_netresult(x->state, x->probs, x->ladder);
        x->node->scoring_cb(x->nodecount, x->state, result, false);
    } else {
        myprintf("leela_b200: asynchronous evaluation failed: %s\n", lb2_last_error());
    }
    delete x;
    opencl.callback_finished();   // the search may only be torn down once this has run (UCTSearch.cpp:737)
}

}  // namespace

// ------------------------------------------------------------------------------------------
// the `opencl` seam
// ------------------------------------------------------------------------------------------
void OpenCL::initialize() {
    if (g_ctx) return;
    if (lb2_init(cfg_gpus.empty() ? nullptr : cfg_gpus.data(), (int)cfg_gpus.size(), &g_ctx)) die("lb2_init");
}
void OpenCL::ensure_thread_initialized() {}
std::string OpenCL::get_device_name() { return g_ctx ? lb2_backend_name(g_ctx) : "B200 evaluator (not initialised)"; }
bool OpenCL::thread_can_issue() { return t_results_outstanding.load(std::memory_order_acquire) < g_max_outstanding; }
std::atomic<int>* OpenCL::get_thread_results_outstanding() { return &t_results_outstanding; }
void OpenCL::callback_started() { m_cb_outstanding.fetch_add(1, std::memory_order_release); }
void OpenCL::callback_finished() { m_cb_outstanding.fetch_sub(1, std::memory_order_release); }
void OpenCL::join_outstanding_cb() {
    if (g_ctx) lb2_drain(g_ctx);
    while (m_cb_outstanding.load(std::memory_order_acquire) > 0) {}
}

// ------------------------------------------------------------------------------------------
// Network
// ------------------------------------------------------------------------------------------
Network* Network::get_Network(void) {
    if (!s_Net) {
        s_Net = new Network();
        s_Net->initialize();
    }
    return s_Net;
}

void Network::initialize(void) {
    myprintf("Initializing B200 evaluator\n");
    opencl.initialize();
    std::string path = g_weights_path;
    if (path.empty()) { const char* e = getenv("LB2_WEIGHTS"); if (e) path = e; }
    if (path.empty()) throw std::runtime_error("leela_b200: no weights file (use --weights or LB2_WEIGHTS)");
    myprintf("Transferring weights to GPU...");
    load_weights(path);
    myprintf("done\n");
}

std::string Network::get_backend() { return opencl.get_device_name(); }

int Network::rotate_nn_idx(const int vertex, int symmetry) {
    assert(vertex >= 0 && vertex < 19 * 19 && symmetry >= 0 && symmetry < 8);
    int x = vertex % 19, y = vertex / 19;
    if (symmetry >= 4) { std::swap(x, y); symmetry -= 4; }
    if (symmetry & 1) y = 18 - y;
    if (symmetry & 2) x = 18 - x;
    return y * 19 + x;
}

int Network::rev_rotate_nn_idx(const int vertex, int symmetry) {
    static const int inverse[8] = {0, 1, 2, 3, 4, 6, 5, 7};
    return rotate_nn_idx(vertex, inverse[symmetry]);
}

void Network::softmax(std::vector<float>& input, std::vector<float>& output, float temperature) {
    assert(&input != &output);
    const float peak = *std::max_element(input.begin(), input.end()) / temperature;
    float denom = 0.0f;
    std::vector<float> e(output.size());
    for (size_t i = 0; i < output.size(); i++) { e[i] = std::exp(input[i] / temperature - peak); denom += e[i]; }
    for (size_t i = 0; i < output.size(); i++) output[i] = e[i] / denom;
}

void Network::gather_features_policy(FastState* state, NNPlanes& planes, BoardPlane** ladder_out) {
    uint32_t packed[361];
    leela_b200::pack_features(state, false, packed, nullptr);
    planes.assign(POLICY_CHANNELS, BoardPlane());
    for (int idx = 0; idx < 361; idx++)
        for (int c = 0; c < POLICY_CHANNELS; c++) planes[c][idx] = (packed[idx] >> c) & 1u;
    if (ladder_out) *ladder_out = &planes[kPolicyPlanes.ladder];
}

void Network::gather_features_value(FastState* state, NNPlanes& planes) {
    uint32_t packed[361];
    leela_b200::pack_features(state, true, packed, nullptr);
    planes.assign(VALUE_CHANNELS, BoardPlane());
    for (int idx = 0; idx < 361; idx++)
        for (int c = 0; c < VALUE_CHANNELS; c++) planes[c][idx] = (packed[idx] >> c) & 1u;
}

Network::Netresult Network::get_scored_moves_internal(FastState* state, NNPlanes& planes, int rotation) {
    uint32_t packed[361];
    for (int idx = 0; idx < 361; idx++) {
        uint32_t w = 0;
        for (size_t c = 0; c < planes.size() && c < 32; c++) w |= (uint32_t)planes[c][idx] << c;
        packed[idx] = w;
    }
    float probs[361];
    const uint8_t rot = (uint8_t)rotation;
    Waiter w;
    if (lb2_submit_policy(g_ctx, packed, &rot, 1, cfg_softmax_temp, probs, Waiter::signal, &w)) die("lb2_submit_policy");
    w.wait("policy evaluation");
    return to_netresult(*state, probs, BoardPlane());
}

float Network::get_value_internal(FastState*, NNPlanes& planes, int rotation) {
    uint32_t packed[361];
    for (int idx = 0; idx < 361; idx++) {
        uint32_t w = 0;
        for (size_t c = 0; c < planes.size() && c < 32; c++) w |= (uint32_t)planes[c][idx] << c;
        packed[idx] = w;
    }
    float win = 0.5f;
    const uint8_t rot = (uint8_t)rotation;
    Waiter w;
    if (lb2_submit_value(g_ctx, packed, &rot, 1, &win, Waiter::signal, &w)) die("lb2_submit_value");
    w.wait("value evaluation");
    return win;
}

Network::Netresult Network::get_scored_moves(FastState* state, Ensemble ensemble, int rotation) {
    Netresult result;
    if (state->board.get_boardsize() != 19) return result;
    get_Network();
    uint32_t one[361];
    BoardPlane ladder;
    leela_b200::pack_features(state, false, one, &ladder);
    if (ensemble == DIRECT) {
        assert(rotation >= 0 && rotation <= 7);
    } else if (ensemble == RANDOM_ROTATION) {
        assert(rotation == -1);
        rotation = Random::get_Rng()->randfix<8>();
    } else {
        assert(ensemble == AVERAGE_ALL);
    }
    float probs[361];
    if (ensemble == AVERAGE_ALL) {
        // the 8 symmetries are expanded, evaluated and averaged (r = 0..7, then / 8) on the device
        if (lb2_eval_ensemble(g_ctx, one, nullptr, 1, cfg_softmax_temp, probs, nullptr)) die("lb2_eval_ensemble");
    } else {
        const uint8_t rot = (uint8_t)rotation;
        Waiter w;
        if (lb2_submit_policy(g_ctx, one, &rot, 1, cfg_softmax_temp, probs, Waiter::signal, &w)) die("lb2_submit_policy");
        w.wait("policy evaluation");
    }
    return to_netresult(*state, probs, ladder);
}

float Network::get_value(FastState* state, Ensemble ensemble) {
    if (state->board.get_boardsize() != 19) {
        assert(false);
        return 0.5f;
    }
    get_Network();
    uint32_t one[361];
    leela_b200::pack_features(state, true, one, nullptr);
    int rotation = 0;
    if (ensemble == RANDOM_ROTATION) rotation = Random::get_Rng()->randfix<8>();
    else assert(ensemble == DIRECT || ensemble == AVERAGE_ALL);
    float win = 0.5f;
    if (ensemble == AVERAGE_ALL) {
        if (lb2_eval_ensemble(g_ctx, nullptr, one, 1, 1.0f, nullptr, &win)) die("lb2_eval_ensemble");
    } else {
        const uint8_t rot = (uint8_t)rotation;
        Waiter w;
        if (lb2_submit_value(g_ctx, one, &rot, 1, &win, Waiter::signal, &w)) die("lb2_submit_value");
        w.wait("value evaluation");
    }
    return win;
}

void Network::async_scored_moves(std::atomic<int>* nodecount, FastState* state, UCTNode* node, Ensemble ensemble, int rotation) {
    if (state->board.get_boardsize() != 19) return;
    assert(ensemble == DIRECT || ensemble == RANDOM_ROTATION);
    if (ensemble == RANDOM_ROTATION) {
        assert(rotation == -1);
        rotation = Random::get_Rng()->randfix<8>();
    }
    AsyncExpansion* x = new AsyncExpansion();
    x->nodecount = nodecount;
    x->state = *state;   // the caller's state moves on; the callback needs its own copy
    x->node = node;
    x->rotation = (uint8_t)rotation;
    x->thread_outstanding = opencl.get_thread_results_outstanding();
    leela_b200::pack_features(state, false, x->planes, &x->ladder);
    x->thread_outstanding->fetch_add(1, std::memory_order_release);
    opencl.callback_started();
    if (lb2_submit_policy(g_ctx, x->planes, &x->rotation, 1, cfg_softmax_temp, x->probs, async_done, x)) {
        x->thread_outstanding->fetch_sub(1, std::memory_order_release);
        opencl.callback_finished();
        delete x;
        die("lb2_submit_policy");
    }
}

// Network::benchmark (Network.cpp:147-199): 2000 policy and 10000 value evaluations spread over the
// search threads, each a single-position request as in the reference; the library batches them.
// Network::benchmark (Network.cpp:147-199): cfg_num_threads threads call get_scored_moves / get_value on copies of the state,
// batch 1 per call. Same legs and the same output lines; the amounts are 50x the reference's 2000 / 10000 (sized for a CPU
// that does ~1 k/s — here they would be over before the reference's 10 ms timer ticks twice) and the clock is
// std::chrono. A third line says what one call costs on the host before anything reaches the device.
void Network::benchmark(FastState* state) {
    const int cpus = cfg_num_threads;
    struct Leg { const char* what; int amount; bool policy; };
    const Leg legs[2] = {{"predictions", 50 * 2000, true}, {"evaluations", 50 * 10000, false}};
    for (const Leg& leg : legs) {
        const int iters_per_thread = (leg.amount + cpus - 1) / cpus;
        const auto start = std::chrono::steady_clock::now();
        ThreadGroup tg(thread_pool);
        for (int i = 0; i < cpus; i++) {
            tg.add_task([iters_per_thread, state, &leg]() {
                FastState mystate = *state;
                for (int loop = 0; loop < iters_per_thread; loop++) {
                    if (leg.policy) get_scored_moves(&mystate, Ensemble::RANDOM_ROTATION);
                    else get_value(&mystate, Ensemble::RANDOM_ROTATION);
                }
            });
        }
        tg.wait_all();
        const float seconds = std::chrono::duration<float>(std::chrono::steady_clock::now() - start).count();
        const int done = iters_per_thread * cpus;
        myprintf("%5d %s in %5.2f seconds -> %d p/s\n", done, leg.what, seconds, (int)((float)done / seconds));
    }
    {
        FastState mystate = *state;
        uint32_t packed[361];
        const auto start = std::chrono::steady_clock::now();
        for (int i = 0; i < 2000; i++) leela_b200::pack_features(&mystate, (i & 1) != 0, packed, nullptr);
        const float us = std::chrono::duration<float>(std::chrono::steady_clock::now() - start).count() * 1e6f / 2000.0f;
        myprintf("feature planes: %.1f us per position on one core (host work in front of every request)\n", us);
    }
}

void Network::show_heatmap(FastState* state, Netresult& result, bool topmoves) {
    float score_at[361];
    std::fill(score_at, score_at + 361, 0.0f);   // non-empty squares are not scored
    for (const scored_node& sn : result) {
        const std::pair<int, int> xy = state->board.get_xy(sn.second);
        if (xy.first >= 0 && xy.first < 19 && xy.second >= 0 && xy.second < 19) score_at[xy.second * 19 + xy.first] = sn.first;
    }
    for (int y = 18; y >= 0; y--) {
        std::string line;
        char cell[16];
        for (int x = 0; x < 19; x++) { snprintf(cell, sizeof cell, "%3d ", int(score_at[y * 19 + x] * 1000)); line += cell; }
        myprintf("%s\n", line.c_str());
    }
    if (!topmoves) return;
    Netresult moves = result;
    std::stable_sort(moves.rbegin(), moves.rend());
    float cum = 0.0f;
    for (size_t tried = 0; cum < 0.85f && tried < moves.size() && moves[tried].first >= 0.01f; tried++) {
        myprintf("%1.3f (%s)\n", moves[tried].first, state->board.move_to_text(moves[tried].second).c_str());
        cum += moves[tried].first;
    }
}

void Network::autotune_from_file(std::string) {
    myprintf("autotune is not part of the evaluation path and is not built into this engine\n");
}
