/* Harness around the UNMODIFIED reference sources (compiled from /root/reference where they
 * lie; see oracle/ref/Makefile). TEST INFRASTRUCTURE ONLY — the product never links this.
 *
 * It drives the reference's own BLAS evaluation path:
 *   - positions   : GameState::init_game(19, 7.5) + play_random_move, as
 *                   Playout::do_playout_benchmark does (Playout.cpp:146-150)
 *   - planes      : Network::gather_features_policy / _value (Network.cpp:883-1201)
 *   - evaluation  : Network::get_scored_moves_internal / get_value_internal
 *                   (Network.cpp:742-832, 676-740) = im2col + cblas_sgemm + bias + ELU,
 *                   softmax(T), (1+tanh)/2
 *   - API level   : Network::get_scored_moves / get_value with AVERAGE_ALL (Network.cpp:590-674)
 *   - single layer: the convolve<> / innerproduct<> templates (Network.cpp:344-423)
 *   - timing      : a Network::benchmark-style loop (Network.cpp:147-199) over given positions
 *
 * The private pieces are reached by including Network.cpp as part of this TU with
 * `private` made public (Network.o is therefore left out of the link).
 */
#include <bits/stdc++.h>
#define private public
#include "Network.cpp"
#undef private

#include "GameState.h"
#include "Zobrist.h"
#include "../../leela_b200/host/b200_network.h"
#include "Matcher.h"
#include "ThreadPool.h"

namespace {

constexpr int P = 361;

struct Positions {
    int n = 0;
    std::vector<uint32_t> pol, val;  // [n][361], bit c = plane c
    std::vector<uint8_t> rot;        // [n]
    std::vector<int32_t> to_move, movenum;
};

void init_reference(int threads) {
    GTP::setup_default_parameters();
    cfg_quiet = true;
    cfg_num_threads = threads;
    thread_pool.initialize(threads);
    auto rng = std::make_unique<Random>(5489);
    Zobrist::init_zobrist(*rng);  // as Leela.cpp:279-280
    Matcher::get_Matcher();
    Network::get_Network();
}

void pack(const Network::NNPlanes& planes, uint32_t* out) {
    for (int i = 0; i < P; i++) {
        uint32_t w = 0;
        for (size_t c = 0; c < planes.size() && c < 32; c++) w |= (uint32_t)planes[c][i] << c;
        out[i] = w;
    }
}

void unpack(const uint32_t* in, Network::NNPlanes& planes) {
    planes.assign(32, Network::BoardPlane());
    for (int i = 0; i < P; i++)
        for (int c = 0; c < 32; c++) planes[c][i] = (in[i] >> c) & 1u;
}

void write_positions(const char* path, const Positions& ps) {
    FILE* f = fopen(path, "wb");
    if (!f) { perror(path); exit(2); }
    int32_t hdr[2] = {ps.n, 0};
    fwrite("LB2POS01", 1, 8, f);
    fwrite(hdr, 4, 2, f);
    fwrite(ps.pol.data(), 4, ps.pol.size(), f);
    fwrite(ps.val.data(), 4, ps.val.size(), f);
    std::vector<uint8_t> r(ps.rot);
    r.resize((ps.n + 3) / 4 * 4, 0);
    fwrite(r.data(), 1, r.size(), f);
    fwrite(ps.to_move.data(), 4, ps.n, f);
    fwrite(ps.movenum.data(), 4, ps.n, f);
    fclose(f);
}

Positions read_positions(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) { perror(path); exit(2); }
    char magic[8];
    int32_t hdr[2];
    Positions ps;
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "LB2POS01", 8) || fread(hdr, 4, 2, f) != 2) {
        fprintf(stderr, "%s: not a positions file\n", path); exit(2);
    }
    ps.n = hdr[0];
    ps.pol.resize((size_t)ps.n * P); ps.val.resize((size_t)ps.n * P);
    ps.rot.resize((ps.n + 3) / 4 * 4); ps.to_move.resize(ps.n); ps.movenum.resize(ps.n);
    bool ok = fread(ps.pol.data(), 4, ps.pol.size(), f) == ps.pol.size()
           && fread(ps.val.data(), 4, ps.val.size(), f) == ps.val.size()
           && fread(ps.rot.data(), 1, ps.rot.size(), f) == ps.rot.size()
           && fread(ps.to_move.data(), 4, ps.n, f) == (size_t)ps.n
           && fread(ps.movenum.data(), 4, ps.n, f) == (size_t)ps.n;
    if (!ok) { fprintf(stderr, "%s: truncated\n", path); exit(2); }
    fclose(f);
    return ps;
}

/* Seeded self-play; keeps a snapshot of roughly every `stride`-th position. Position 0 is
 * the empty board (the reference's self-test position, GTP.cpp:105-125). */
Positions generate(int n, uint32_t seed, const std::function<void(FastState&, int)>& hook) {
    Positions ps;
    ps.n = n;
    ps.pol.resize((size_t)n * P); ps.val.resize((size_t)n * P);
    ps.rot.resize(n); ps.to_move.resize(n); ps.movenum.resize(n);
    Random::get_Rng()->seedrandom(seed);
    std::mt19937 pick(seed * 2654435761u + 17u);
    int got = 0;
    auto snapshot = [&](GameState& g) {
        FastState s = g;
        Network::NNPlanes pp, vp;
        Network::BoardPlane* ladder = nullptr;
        Network::gather_features_policy(&s, pp, &ladder);
        Network::gather_features_value(&s, vp);
        pack(pp, &ps.pol[(size_t)got * P]);
        pack(vp, &ps.val[(size_t)got * P]);
        ps.rot[got] = (uint8_t)(got % 8);
        ps.to_move[got] = s.get_to_move();
        ps.movenum[got] = s.get_movenum();
        if (hook) hook(s, got);
        got++;
    };
    GameState game;
    game.init_game(19, 7.5f);
    snapshot(game);
    const int playoutlen = 19 * 19 * 2;
    const int resign = (19 * 19) / 3;
    while (got < n) {
        game.init_game(19, 7.5f);
        int next = 1 + (int)(pick() % 24);
        do {
            game.play_random_move(game.get_to_move());
            if ((int)game.get_movenum() >= next && got < n) {
                snapshot(game);
                next = game.get_movenum() + 1 + (int)(pick() % 24);
            }
        } while (got < n && game.get_passes() < 2 && (int)game.get_movenum() < playoutlen
                 && abs(game.estimate_mc_score()) < resign);
    }
    return ps;
}

/* Full 361-vector policy for arbitrary planes: get_scored_moves_internal only consults the
 * state for its EMPTY filter (Network.cpp:820-829), so an empty board keeps every point. */
void eval_direct(const Positions& ps, int i, FastState& empty_state, float* pol_out, float* val_out) {
    Network::NNPlanes pp, vp;
    unpack(&ps.pol[(size_t)i * P], pp);
    unpack(&ps.val[(size_t)i * P], vp);
    auto res = Network::get_scored_moves_internal(&empty_state, pp, ps.rot[i]);
    if ((int)res.size() != P) { fprintf(stderr, "unexpected result size %zu\n", res.size()); exit(3); }
    for (int idx = 0; idx < P; idx++) {
        int vtx = empty_state.board.get_vertex(idx % 19, idx / 19);
        if (res[idx].second != vtx) { fprintf(stderr, "vertex order mismatch\n"); exit(3); }
        pol_out[idx] = res[idx].first;
    }
    *val_out = Network::get_value_internal(&empty_state, vp, ps.rot[i]);
}

int cmd_gen(int argc, char** argv, bool eval) {
    if (argc < 5) { fprintf(stderr, "usage: %s <out> <n> <seed> [n_avg]\n", argv[1]); return 2; }
    const char* out = argv[2];
    int n = atoi(argv[3]);
    uint32_t seed = (uint32_t)strtoul(argv[4], nullptr, 10);
    int n_avg = argc > 5 ? atoi(argv[5]) : 0;
    init_reference(1);
    /* AVERAGE_ALL at API level needs the real state (EMPTY filter + ladder prune,
     * Network.cpp:643-667), so it is evaluated while the positions are generated. */
    std::vector<float> pol_avg((size_t)n_avg * P, -1.0f), val_avg(n_avg, 0.0f);
    auto hook = [&](FastState& s, int i) {
        if (!eval || i >= n_avg) return;
        auto res = Network::get_scored_moves(&s, Network::AVERAGE_ALL);
        for (auto& sn : res) {
            auto xy = s.board.get_xy(sn.second);
            pol_avg[(size_t)i * P + xy.second * 19 + xy.first] = sn.first;
        }
        val_avg[i] = Network::get_value(&s, Network::AVERAGE_ALL);
    };
    Positions ps = generate(n, seed, hook);
    std::string base(out);
    write_positions((base + ".pos").c_str(), ps);
    if (!eval) return 0;

    FastState empty_state;
    empty_state.init_game(19, 7.5f);
    std::vector<float> pol((size_t)n * P), val(n);
    for (int i = 0; i < n; i++) eval_direct(ps, i, empty_state, &pol[(size_t)i * P], &val[i]);
    FILE* f = fopen((base + ".out").c_str(), "wb");
    if (!f) { perror(out); return 2; }
    int32_t hdr[2] = {n, n_avg};
    float temp = cfg_softmax_temp;
    fwrite("LB2OUT01", 1, 8, f);
    fwrite(hdr, 4, 2, f);
    fwrite(&temp, 4, 1, f);
    fwrite(pol.data(), 4, pol.size(), f);
    fwrite(val.data(), 4, val.size(), f);
    fwrite(pol_avg.data(), 4, pol_avg.size(), f);
    fwrite(val_avg.data(), 4, val_avg.size(), f);
    fclose(f);
    return 0;
}

/* Evaluate given planes + rotations (any bit patterns, e.g. hand-made edge cases). */
int cmd_eval(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: eval <in.pos> <out>\n"); return 2; }
    init_reference(1);
    Positions ps = read_positions(argv[2]);
    FastState empty_state;
    empty_state.init_game(19, 7.5f);
    std::vector<float> pol((size_t)ps.n * P), val(ps.n);
    for (int i = 0; i < ps.n; i++) eval_direct(ps, i, empty_state, &pol[(size_t)i * P], &val[i]);
    FILE* f = fopen(argv[3], "wb");
    if (!f) { perror(argv[3]); return 2; }
    int32_t hdr[2] = {ps.n, 0};
    float temp = cfg_softmax_temp;
    fwrite("LB2OUT01", 1, 8, f);
    fwrite(hdr, 4, 2, f);
    fwrite(&temp, 4, 1, f);
    fwrite(pol.data(), 4, pol.size(), f);
    fwrite(val.data(), 4, val.size(), f);
    fclose(f);
    return 0;
}

std::vector<float> read_floats(const char* path, size_t n) {
    std::vector<float> v(n);
    FILE* f = fopen(path, "rb");
    if (!f || fread(v.data(), 4, n, f) != n) { fprintf(stderr, "%s: short read (want %zu floats)\n", path, n); exit(2); }
    fclose(f);
    return v;
}

template <unsigned K, unsigned C, unsigned O>
void run_conv(char** a) {
    constexpr size_t WN = (size_t)K * K * C * O;
    auto in = read_floats(a[0], (size_t)C * P);
    auto wv = read_floats(a[1], WN);
    auto bv = read_floats(a[2], O);
    auto w = std::make_unique<std::array<float, WN>>();
    auto b = std::make_unique<std::array<float, O>>();
    std::copy(wv.begin(), wv.end(), w->begin());
    std::copy(bv.begin(), bv.end(), b->begin());
    std::vector<float> out((size_t)O * P);
    convolve<K, C, O>(in, *w, *b, out);
    FILE* f = fopen(a[3], "wb");
    fwrite(out.data(), 4, out.size(), f);
    fclose(f);
}

template <unsigned I, unsigned O>
void run_ip(char** a) {
    auto in = read_floats(a[0], I);
    auto wv = read_floats(a[1], (size_t)I * O);
    auto bv = read_floats(a[2], O);
    auto w = std::make_unique<std::array<float, (size_t)I * O>>();
    auto b = std::make_unique<std::array<float, O>>();
    std::copy(wv.begin(), wv.end(), w->begin());
    std::copy(bv.begin(), bv.end(), b->begin());
    std::vector<float> out(O);
    innerproduct<I, O>(in, *w, *b, out);
    FILE* f = fopen(a[3], "wb");
    fwrite(out.data(), 4, out.size(), f);
    fclose(f);
}

/* One layer through the reference's own templates, on caller-supplied tensors. */
int cmd_layer(int argc, char** argv) {
    if (argc < 9) { fprintf(stderr, "usage: layer conv|ip <k> <cin> <cout> <in> <w> <b> <out>\n"); return 2; }
    openblas_set_num_threads(1);
    std::string kind = argv[2];
    int k = atoi(argv[3]), ci = atoi(argv[4]), co = atoi(argv[5]);
    char** files = argv + 6;
    if (kind == "conv") {
        if (k == 5 && ci == 32 && co == 96) run_conv<5, 32, 96>(files);
        else if (k == 3 && ci == 96 && co == 128) run_conv<3, 96, 128>(files);
        else if (k == 3 && ci == 128 && co == 128) run_conv<3, 128, 128>(files);
        else if (k == 3 && ci == 128 && co == 1) run_conv<3, 128, 1>(files);
        else if (k == 5 && ci == 32 && co == 64) run_conv<5, 32, 64>(files);
        else if (k == 3 && ci == 64 && co == 64) run_conv<3, 64, 64>(files);
        else if (k == 3 && ci == 64 && co == 1) run_conv<3, 64, 1>(files);
        else { fprintf(stderr, "shape not in the reference nets\n"); return 2; }
    } else {
        if (ci == 361 && co == 256) run_ip<361, 256>(files);
        else if (ci == 256 && co == 1) run_ip<256, 1>(files);
        else { fprintf(stderr, "shape not in the reference nets\n"); return 2; }
    }
    return 0;
}

/* Network::benchmark-style throughput (Network.cpp:147-199): `threads` pool threads each loop
 * single-position forwards, OpenBLAS pinned to one thread per worker (Network.cpp:239), for
 * about `seconds`. which = policy | value | both. Prints one JSON line. */
int cmd_bench(int argc, char** argv) {
    if (argc < 6) { fprintf(stderr, "usage: bench <in.pos> <threads> <seconds> policy|value|both\n"); return 2; }
    Positions ps = read_positions(argv[2]);
    int threads = atoi(argv[3]);
    double seconds = atof(argv[4]);
    std::string which = argv[5];
    init_reference(threads);
    std::atomic<long> done{0};
    std::atomic<bool> stop{false};
    auto t0 = std::chrono::steady_clock::now();
    ThreadGroup tg(thread_pool);
    for (int t = 0; t < threads; t++) {
        tg.add_task([&, t]() {
            FastState empty_state;
            empty_state.init_game(19, 7.5f);
            Network::NNPlanes pp, vp;
            int i = t % ps.n;
            volatile float sink = 0;
            while (!stop.load(std::memory_order_relaxed)) {
                if (which != "value") {
                    unpack(&ps.pol[(size_t)i * P], pp);
                    auto r = Network::get_scored_moves_internal(&empty_state, pp, ps.rot[i]);
                    sink = sink + r[0].first;
                }
                if (which != "policy") {
                    unpack(&ps.val[(size_t)i * P], vp);
                    sink = sink + Network::get_value_internal(&empty_state, vp, ps.rot[i]);
                }
                done.fetch_add(1, std::memory_order_relaxed);
                i = (i + threads) % ps.n;
                if (t == 0) {
                    std::chrono::duration<double> el = std::chrono::steady_clock::now() - t0;
                    if (el.count() >= seconds) stop.store(true);
                }
            }
        });
    }
    tg.wait_all();
    std::chrono::duration<double> el = std::chrono::steady_clock::now() - t0;
    printf("{\"positions\": %ld, \"seconds\": %.4f, \"pos_per_s\": %.3f, \"threads\": %d, \"which\": \"%s\", \"blas_core\": \"%s\"}\n",
           done.load(), el.count(), done.load() / el.count(), threads, which.c_str(), openblas_get_corename());
    return 0;
}

/* Fixed-work timing for bench.py --impl reference: `warmup` untimed then `steps` timed steps,
 * each step = `sample` positions (policy and/or value forwards, batch 1 each, like
 * Network::benchmark) shared over `threads` pool threads. Prints one JSON line. */
int cmd_steps(int argc, char** argv) {
    if (argc < 8) { fprintf(stderr, "usage: steps <in.pos> <threads> <sample> <warmup> <steps> policy|value|both\n"); return 2; }
    Positions ps = read_positions(argv[2]);
    int threads = atoi(argv[3]), sample = atoi(argv[4]), warmup = atoi(argv[5]), steps = atoi(argv[6]);
    std::string which = argv[7];
    init_reference(threads);
    double timed = 0.0;
    for (int s = 0; s < warmup + steps; s++) {
        std::atomic<int> next{0};
        auto t0 = std::chrono::steady_clock::now();
        ThreadGroup tg(thread_pool);
        for (int t = 0; t < threads; t++) {
            tg.add_task([&, s]() {
                FastState empty_state;
                empty_state.init_game(19, 7.5f);
                Network::NNPlanes pp, vp;
                volatile float sink = 0;
                for (;;) {
                    int k = next.fetch_add(1);
                    if (k >= sample) break;
                    int i = (s * sample + k) % ps.n;
                    if (which != "value") {
                        unpack(&ps.pol[(size_t)i * P], pp);
                        auto r = Network::get_scored_moves_internal(&empty_state, pp, ps.rot[i]);
                        sink = sink + r[0].first;
                    }
                    if (which != "policy") {
                        unpack(&ps.val[(size_t)i * P], vp);
                        sink = sink + Network::get_value_internal(&empty_state, vp, ps.rot[i]);
                    }
                }
            });
        }
        tg.wait_all();
        std::chrono::duration<double> el = std::chrono::steady_clock::now() - t0;
        if (s >= warmup) timed += el.count();
    }
    printf("{\"positions\": %ld, \"seconds\": %.6f, \"pos_per_s\": %.3f, \"threads\": %d, \"sample\": %d, \"steps\": %d, \"which\": \"%s\", \"blas_core\": \"%s\"}\n",
           (long)sample * steps, timed, (double)sample * steps / timed, threads, sample, steps, which.c_str(), openblas_get_corename());
    return 0;
}

/* Drop-in check at the API level: the UNMODIFIED reference engine supplies live game states,
 * feature planes (gather_features_*) and its own answers (get_scored_moves / get_value through
 * the public API, all three ensembles); the B200 evaluator answers the same calls through the
 * host mirror leela_b200/host/b200_network.h on top of the C ABI. Needs a B200. Prints JSON. */
int cmd_apicheck(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: apicheck <n> <seed>\n"); return 2; }
    const int n = atoi(argv[2]);
    const uint32_t seed = (uint32_t)strtoul(argv[3], nullptr, 10);
    init_reference(1);
    leela_b200::B200Network net;
    try {
        net.initialize();
        net.push_convolve(LB2_POLICY, 5, conv1_w, conv1_b);   net.push_convolve(LB2_POLICY, 3, conv2_w, conv2_b);
        net.push_convolve(LB2_POLICY, 3, conv3_w, conv3_b);   net.push_convolve(LB2_POLICY, 3, conv4_w, conv4_b);
        net.push_convolve(LB2_POLICY, 3, conv5_w, conv5_b);   net.push_convolve(LB2_POLICY, 3, conv6_w, conv6_b);
        net.push_convolve(LB2_POLICY, 3, conv7_w, conv7_b);   net.push_convolve(LB2_POLICY, 3, conv8_w, conv8_b);
        net.push_convolve(LB2_POLICY, 3, conv9_w, conv9_b);   net.push_convolve(LB2_POLICY, 3, conv10_w, conv10_b);
        net.push_convolve(LB2_POLICY, 3, conv11_w, conv11_b); net.push_convolve(LB2_POLICY, 3, conv12_w, conv12_b);
        net.push_convolve(LB2_POLICY, 3, conv13_w, conv13_b);
        net.push_convolve(LB2_VALUE, 5, val_conv1_w, val_conv1_b);   net.push_convolve(LB2_VALUE, 3, val_conv2_w, val_conv2_b);
        net.push_convolve(LB2_VALUE, 3, val_conv3_w, val_conv3_b);   net.push_convolve(LB2_VALUE, 3, val_conv4_w, val_conv4_b);
        net.push_convolve(LB2_VALUE, 3, val_conv5_w, val_conv5_b);   net.push_convolve(LB2_VALUE, 3, val_conv6_w, val_conv6_b);
        net.push_convolve(LB2_VALUE, 3, val_conv7_w, val_conv7_b);   net.push_convolve(LB2_VALUE, 3, val_conv8_w, val_conv8_b);
        net.push_convolve(LB2_VALUE, 3, val_conv9_w, val_conv9_b);   net.push_convolve(LB2_VALUE, 3, val_conv10_w, val_conv10_b);
        net.push_convolve(LB2_VALUE, 3, val_conv11_w, val_conv11_b); net.push_convolve(LB2_VALUE, 3, val_conv12_w, val_conv12_b);
        net.push_innerproduct(LB2_VALUE, val_ip13_w, val_ip13_b);
        net.push_innerproduct(LB2_VALUE, val_ip14_w, val_ip14_b);
        net.finalize();
    } catch (const std::exception& e) {
        printf("{\"error\": \"%s\"}\n", e.what());
        return 1;
    }
    double max_dp = 0, max_dv = 0, max_dp_avg = 0, max_dv_avg = 0;
    int top1_same = 0, order_mismatch = 0, cases = 0, ladder_zeroed = 0;
    auto rng_fixed = [](int r) { return [r]() { return r; }; };
    auto hook = [&](FastState& s, int i) {
        Network::NNPlanes pp, vp;
        Network::BoardPlane* ladder = nullptr;
        Network::gather_features_policy(&s, pp, &ladder);
        Network::gather_features_value(&s, vp);
        const int r = i % 8;
        auto compare = [&](const Network::Netresult& a, const leela_b200::B200Network::Netresult& b, double& worst) {
            if (a.size() != b.size()) { order_mismatch++; return; }
            size_t ba = 0, bb = 0;
            for (size_t k = 0; k < a.size(); k++) {
                if (a[k].second != b[k].second) { order_mismatch++; return; }
                worst = std::max(worst, (double)std::fabs(a[k].first - b[k].first));
                if (a[k].first > a[ba].first) ba = k;
                if (b[k].first > b[bb].first) bb = k;
                if (a[k].first == 0.0f && b[k].first == 0.0f) ladder_zeroed++;
            }
            if (a.empty() || ba == bb || std::fabs(a[ba].first - a[bb].first) < 6e-3) top1_same++;
        };
        auto ref_d = Network::get_scored_moves(&s, Network::DIRECT, r);
        auto our_d = net.get_scored_moves(&s, pp, ladder, leela_b200::B200Network::DIRECT, r, cfg_softmax_temp, rng_fixed(0));
        compare(ref_d, our_d, max_dp);
        max_dv = std::max(max_dv, (double)std::fabs(Network::get_value(&s, Network::DIRECT) -
                                                    net.get_value(&s, vp, leela_b200::B200Network::DIRECT, rng_fixed(0))));
        if (i % 4 == 0) {
            auto ref_a = Network::get_scored_moves(&s, Network::AVERAGE_ALL);
            auto our_a = net.get_scored_moves(&s, pp, ladder, leela_b200::B200Network::AVERAGE_ALL, -1, cfg_softmax_temp, rng_fixed(0));
            compare(ref_a, our_a, max_dp_avg);
            max_dv_avg = std::max(max_dv_avg, (double)std::fabs(Network::get_value(&s, Network::AVERAGE_ALL) -
                                                                net.get_value(&s, vp, leela_b200::B200Network::AVERAGE_ALL, rng_fixed(0))));
            cases++;
        }
        cases++;
    };
    generate(n, seed, hook);
    printf("{\"positions\": %d, \"cases\": %d, \"max_dp_direct\": %.6g, \"max_dv_direct\": %.6g, \"max_dp_average_all\": %.6g, "
           "\"max_dv_average_all\": %.6g, \"top1_agree_or_near_tie\": %d, \"order_mismatch\": %d, \"ladder_zeroed_points\": %d, "
           "\"backend\": \"%s\"}\n",
           n, cases, max_dp, max_dv, max_dp_avg, max_dv_avg, top1_same, order_mismatch, ladder_zeroed, net.get_backend().c_str());
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: ref_harness dump|planes|eval|layer|bench|steps ...\n");
        return 2;
    }
    std::string cmd = argv[1];
    if (cmd == "dump") return cmd_gen(argc, argv, true);
    if (cmd == "planes") return cmd_gen(argc, argv, false);
    if (cmd == "eval") return cmd_eval(argc, argv);
    if (cmd == "layer") return cmd_layer(argc, argv);
    if (cmd == "bench") return cmd_bench(argc, argv);
    if (cmd == "steps") return cmd_steps(argc, argv);
    if (cmd == "apicheck") return cmd_apicheck(argc, argv);
    fprintf(stderr, "unknown command %s\n", cmd.c_str());
    return 2;
}
