// b200_network.h — host-side C++ mirror of the reference's Network scoring API on top of the
// C ABI (include/leela_b200.h).
//
// Same names, argument meaning and error behaviour as Network.h:25-71 / Network.cpp:590-674 for
// the part of the path that lives above the backend seam:
//   Ensemble {DIRECT, RANDOM_ROTATION, AVERAGE_ALL}, scored_node = pair<float prob, int vertex>,
//   Netresult in raster order over EMPTY points only (no renormalisation, Network.cpp:820-829),
//   AVERAGE_ALL = mean over the 8 symmetries (Network.cpp:643-654, 605-615; on the device:
//   lb2_eval_ensemble), losing-ladder points
//   zeroed (Network.cpp:656-667), board != 19 -> empty result / 0.5 (Network.cpp:591-594, 627-629),
//   DIRECT needs rotation 0..7, RANDOM_ROTATION needs -1 (asserts at Network.cpp:636-640).
//
// It is a template over the state type so that it compiles unchanged against the reference's
// FastState (state->board.get_boardsize / get_vertex / get_square / get_xy, FastBoard::EMPTY) —
// see oracle/ref/ref_harness.cpp "apicheck", which links the unmodified reference with this
// header — and against a stand-in board in unit tests. Feature planes come from the reference's
// gather_features_policy / _value (they need FastBoard and stay on the host).
#pragma once
#include <array>
#include <bitset>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/leela_b200.h"

namespace leela_b200 {

class B200Network {
public:
    enum Ensemble { DIRECT, RANDOM_ROTATION, AVERAGE_ALL };
    using BoardPlane = std::bitset<19 * 19>;
    using NNPlanes = std::vector<BoardPlane>;
    using scored_node = std::pair<float, int>;
    using Netresult = std::vector<scored_node>;

    B200Network() = default;
    B200Network(const B200Network&) = delete;
    B200Network& operator=(const B200Network&) = delete;
    ~B200Network() { if (m_ctx) lb2_destroy(m_ctx); }

    // Network::initialize (Network.cpp:201-235): create the context, then push the layers.
    void initialize(const std::vector<int>& gpus = {}) {
        check(lb2_init(gpus.empty() ? nullptr : gpus.data(), (int)gpus.size(), &m_ctx));
        check(lb2_net_create(m_ctx, LB2_POLICY, &m_policy));
        check(lb2_net_create(m_ctx, LB2_VALUE, &m_value));
    }
    // OpenCL_Network::push_convolve semantics (OpenCL.h:66-87): geometry is derived from the array sizes.
    template <class W, class B>
    void push_convolve(int kind, int k, const W& w, const B& b) {
        const int c_out = (int)b.size(), c_in = (int)(w.size() / (size_t)(c_out * k * k));
        check(lb2_net_push_conv(net(kind), k, c_in, c_out, w.data(), b.data()));
    }
    template <class W, class B>
    void push_innerproduct(int kind, const W& w, const B& b) {
        const int n_out = (int)b.size(), n_in = (int)(w.size() / (size_t)n_out);
        check(lb2_net_push_ip(net(kind), n_in, n_out, w.data(), b.data()));
    }
    void finalize() { check(lb2_net_finalize(m_policy)); check(lb2_net_finalize(m_value)); }
    std::string get_backend() const { return lb2_backend_name(m_ctx); }
    lb2_ctx* ctx() const { return m_ctx; }

    // Network::rotate_nn_idx / rev_rotate_nn_idx (Network.cpp:1341-1379)
    static int rotate_nn_idx(int vertex, int symmetry) {
        int x = vertex % 19, y = vertex / 19;
        if (symmetry >= 4) { std::swap(x, y); symmetry -= 4; }
        if (symmetry & 1) y = 18 - y;
        if (symmetry & 2) x = 18 - x;
        return y * 19 + x;
    }
    static int rev_rotate_nn_idx(int vertex, int symmetry) {
        static const int invert[8] = {0, 1, 2, 3, 4, 6, 5, 7};
        return rotate_nn_idx(vertex, invert[symmetry]);
    }

    // NNPlanes (32 x bitset<361>) -> one uint32 per board point, bit c = plane c
    static void pack_planes(const NNPlanes& planes, uint32_t* out) {
        for (int i = 0; i < 361; i++) {
            uint32_t w = 0;
            for (size_t c = 0; c < planes.size() && c < 32; c++) w |= (uint32_t)planes[c][i] << c;
            out[i] = w;
        }
    }

    // Network::get_scored_moves (Network.cpp:624-674) given the already gathered policy planes
    // (`ladder` = plane 25, as gather_features_policy returns it). `rng8` supplies the symmetry for
    // RANDOM_ROTATION (the reference uses Random::get_Rng()->randfix<8>()).
    template <class State, class Rng8>
    Netresult get_scored_moves(State* state, const NNPlanes& planes, const BoardPlane* ladder, Ensemble ensemble,
                               int rotation, float softmax_temp, Rng8&& rng8) {
        Netresult result;
        if (state->board.get_boardsize() != 19) return result;
        uint32_t packed[361];
        float probs[361];
        pack_planes(planes, packed);
        if (ensemble == AVERAGE_ALL) {  // the 8 symmetries are expanded, evaluated and averaged on the device
            check(lb2_eval_ensemble(m_ctx, packed, nullptr, 1, softmax_temp, probs, nullptr));
        } else {
            uint8_t rot;
            if (ensemble == DIRECT) {
                if (rotation < 0 || rotation > 7) throw std::invalid_argument("DIRECT needs rotation 0..7");
                rot = (uint8_t)rotation;
            } else {
                if (rotation != -1) throw std::invalid_argument("RANDOM_ROTATION needs rotation -1");
                rot = (uint8_t)(rng8() & 7);
            }
            check(lb2_eval_policy(m_ctx, packed, &rot, 1, softmax_temp, probs));
        }
        for (int idx = 0; idx < 361; idx++) {
            const int vtx = state->board.get_vertex(idx % 19, idx / 19);
            using Board = typename std::remove_reference<decltype(state->board)>::type;
            if (state->board.get_square(vtx) != Board::EMPTY) continue;
            result.emplace_back(probs[idx], vtx);   // outputs are already un-rotated
        }
        if (ladder) {  // prune losing ladders completely (Network.cpp:656-667)
            for (auto& sm : result) {
                const std::pair<int, int> xy = state->board.get_xy(sm.second);
                if ((*ladder)[xy.second * 19 + xy.first]) sm.first = 0.0f;
            }
        }
        return result;
    }

    // Network::get_value (Network.cpp:590-622) given the value planes.
    template <class State, class Rng8>
    float get_value(State* state, const NNPlanes& planes, Ensemble ensemble, Rng8&& rng8) {
        if (state->board.get_boardsize() != 19) return 0.5f;
        uint32_t packed[361];
        float win = 0.5f;
        pack_planes(planes, packed);
        if (ensemble == AVERAGE_ALL) {
            check(lb2_eval_ensemble(m_ctx, nullptr, packed, 1, 1.0f, nullptr, &win));
        } else {
            const uint8_t rot = ensemble == RANDOM_ROTATION ? (uint8_t)(rng8() & 7) : (uint8_t)0;
            check(lb2_eval_value(m_ctx, packed, &rot, 1, &win));
        }
        return win;
    }

private:
    lb2_net* net(int kind) const { return kind == LB2_POLICY ? m_policy : m_value; }
    static void check(int rc) {
        if (rc != LB2_OK) throw std::runtime_error(std::string("leela_b200: ") + lb2_last_error());  // as OpenCL.cpp:665-669
    }
    lb2_ctx* m_ctx = nullptr;
    lb2_net* m_policy = nullptr;
    lb2_net* m_value = nullptr;
};

}  // namespace leela_b200
