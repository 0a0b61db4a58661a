// main.cpp — entry point of the drop-in engine: the reference's GTP engine (search, board, book,
// time control: its own unmodified sources) with the B200 evaluator behind Network.
//
// Replaces Leela.cpp, whose option parser needs boost::program_options (absent from this image);
// the options keep the reference's names and meaning (Leela.cpp:30-75, 111-224), start-up order
// follows Leela.cpp:253-300. Differences, all on purpose:
//   --threads may exceed the core count (up to MAX_CPUS): search threads mostly wait for the GPU,
//             and the number of threads in flight is what fills a device batch
//   --weights FILE         network weights ("LB2WGT01", leela_b200/fileio.py); or LB2_WEIGHTS
//   --max-outstanding N    async policy requests per search thread (reference: 2, OpenCL.cpp:453)
//   --batch N              positions per device pass (lb2 option max_batch)
//   --precise              lb2 option precise (split-operand evaluation)
//   --own-planes / --check-planes   feature planes through the library's own board, or both ways with a
//                              comparison on every position the search evaluates
//   --dump-planes OUT N SEED   write the feature planes of N seeded self-play positions and exit
//                              (no GPU needed; tests compare them with the reference's own)
//   -DLB2_REFERENCE_BUILD  the same main for the reference's own CPU engine (oracle/ref/Makefile
//                          `ref_engine`, the baseline of `bench.py --engine`): B200 options dropped
//   the OpenCL self-test (GTP.cpp:105-125) pins the reference's 192-wide weights, which are not in the
//   snapshot; in its place leela_b200::self_test checks the known answers the weights file carries
//   (the reference's CPU outputs for those weights on the empty board); --noselftest skips it.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include "config.h"
#include "AttribScores.h"
#include "GameState.h"
#include "GTP.h"
#include "Matcher.h"
#include "Network.h"
#include "Random.h"
#include "ThreadPool.h"
#include "Utils.h"
#include "Zobrist.h"

#ifndef LB2_REFERENCE_BUILD
#include "network_b200.h"
#endif

using namespace Utils;

namespace {

void usage() {
    std::cout << "Allowed options:\n"
                 "  -h, --help              Show commandline options.\n"
                 "  -g, --gtp               Enable GTP mode.\n"
                 "  -t, --threads N         Number of search threads (may exceed the core count).\n"
                 "  -p, --playouts N        Limit the number of playouts. Requires --noponder.\n"
                 "  -b, --lagbuffer CS      Safety margin for time usage in centiseconds.\n"
                 "  -l, --logfile FILE      File to log input/output to.\n"
                 "  -q, --quiet             Disable all diagnostic output.\n"
                 "  -k, --komiadjust        Adjust komi one point in my disadvantage.\n"
                 "      --noponder          Disable thinking on opponent's time.\n"
                 "      --nonets            Disable use of neural networks.\n"
                 "      --nobook            Disable use of the fuseki library.\n"
                 "      --mature_threshold N  Visits before a node's children get policy priors (GPU default 25).\n"
                 "      --eval_thresh N     Visits before a node is sent to the value net (GPU default 2).\n"
                 "      --extra_symmetry N  Visits per extra symmetry of the policy net (GPU default 350).\n"
                 "      --gpu ID            B200 device(s) to use (repeatable; weights replicated, batches sharded).\n"
                 "      --weights FILE      Network weights file (or LB2_WEIGHTS).\n"
                 "      --own-planes        Build the feature planes with the library's own board (lb2_planes_from_position): the default.\n"
                 "      --ref-planes        Build them through the reference's board queries instead (bit-identical, slower).\n"
                 "      --check-planes      Build them both ways and abort on the first difference.\n"
                 "      --max-outstanding N Async policy requests per search thread (default 2).\n"
                 "      --batch N           Positions per device pass (default 256).\n"
                 "      --queue-linger 0|1  lb2 option queue_linger (default 1: requests accumulate while the device computes).\n"
                 "      --precise           Split-operand evaluation (fp16 hi + lo): within 1e-4 of the fp32 nets, ~0.4x throughput.\n"
                 "      --dump-planes OUT N SEED  Dump feature planes of seeded self-play positions and exit.\n";
}

#ifndef LB2_REFERENCE_BUILD
// Seeded self-play snapshots, the same walk as the reference harness (oracle/ref/ref_harness.cpp
// `gen`): position 0 is the empty board, then roughly every 1..24th position of random games.
int dump_planes(const char* out, int n, uint32_t seed) {
    std::vector<uint32_t> pol((size_t)n * 361), val((size_t)n * 361);
    std::vector<uint8_t> rot((n + 3) / 4 * 4, 0);
    std::vector<int32_t> to_move(n), movenum(n);
    // the raw positions beside the planes (OUT.raw, "LB2RAW01"): what lb2_planes_from_position takes
    std::vector<uint8_t> stones((size_t)n * 361);
    std::vector<int32_t> ko(n), last(n), prev(n);
    std::vector<float> komi(n);
    Random::get_Rng()->seedrandom(seed);
    std::mt19937 pick(seed * 2654435761u + 17u);
    int got = 0;
    auto snapshot = [&](GameState& g) {
        FastState s = g;
        leela_b200::pack_features(&s, false, &pol[(size_t)got * 361], nullptr);
        leela_b200::pack_features(&s, true, &val[(size_t)got * 361], nullptr);
        rot[got] = (uint8_t)(got % 8);
        to_move[got] = s.get_to_move();
        movenum[got] = s.get_movenum();
        auto index_of = [&](int vertex) {
            if (vertex <= 0) return -1;   // none or pass
            const std::pair<int, int> xy = s.board.get_xy(vertex);
            return xy.second * 19 + xy.first;
        };
        for (int idx = 0; idx < 361; idx++) {
            const FastBoard::square_t sq = s.board.get_square(s.board.get_vertex(idx % 19, idx / 19));
            stones[(size_t)got * 361 + idx] = sq == FastBoard::BLACK ? 1 : (sq == FastBoard::WHITE ? 2 : 0);
        }
        ko[got] = index_of(s.get_komove());
        last[got] = index_of(s.get_last_move());
        prev[got] = index_of(s.get_prevlast_move());
        komi[got] = s.get_komi();
        got++;
    };
    GameState game;
    game.init_game(19, 7.5f);
    snapshot(game);
    while (got < n) {
        game.init_game(19, 7.5f);
        int next = 1 + (int)(pick() % 24);
        do {
            game.play_random_move(game.get_to_move());
            if ((int)game.get_movenum() >= next && got < n) {
                snapshot(game);
                next = game.get_movenum() + 1 + (int)(pick() % 24);
            }
        } while (got < n && game.get_passes() < 2 && (int)game.get_movenum() < 19 * 19 * 2 &&
                 abs(game.estimate_mc_score()) < (19 * 19) / 3);
    }
    FILE* f = fopen(out, "wb");
    if (!f) { perror(out); return 2; }
    const int32_t hdr[2] = {n, 0};
    fwrite("LB2POS01", 1, 8, f);
    fwrite(hdr, 4, 2, f);
    fwrite(pol.data(), 4, pol.size(), f);
    fwrite(val.data(), 4, val.size(), f);
    fwrite(rot.data(), 1, rot.size(), f);
    fwrite(to_move.data(), 4, n, f);
    fwrite(movenum.data(), 4, n, f);
    fclose(f);
    const std::string raw_path = std::string(out) + ".raw";
    f = fopen(raw_path.c_str(), "wb");
    if (!f) { perror(raw_path.c_str()); return 2; }
    fwrite("LB2RAW01", 1, 8, f);
    fwrite(hdr, 4, 2, f);
    fwrite(stones.data(), 1, stones.size(), f);
    fwrite(to_move.data(), 4, n, f);
    fwrite(ko.data(), 4, n, f);
    fwrite(last.data(), 4, n, f);
    fwrite(prev.data(), 4, n, f);
    fwrite(komi.data(), 4, n, f);
    fclose(f);
    return 0;
}

// printed at exit (GTP `quit` calls exit()): how well the search filled the device
void print_evaluator_stats() {
    if (lb2_ctx* ctx = leela_b200::context()) {
        const long pos = lb2_get_option(ctx, "stat_positions"), bat = lb2_get_option(ctx, "stat_batches");
        fprintf(stderr, "B200 evaluator: %ld positions in %ld device batches (mean batch %.1f), %ld requests\n", pos, bat,
                bat ? (double)pos / bat : 0.0, lb2_get_option(ctx, "stat_requests"));
        if (leela_b200::planes_checked())
            fprintf(stderr, "feature planes cross-checked (own board vs reference board queries): %ld, all identical\n", leela_b200::planes_checked());
    }
}
#endif

}  // namespace

int main(int argc, char* argv[]) {
    bool gtp_mode = false, noponder = false, playouts_set = false;
    long batch = 0, linger = -1;
    bool precise = false, no_selftest = false;
    const char* dump_out = nullptr;
    int dump_n = 0;
    uint32_t dump_seed = 0;

    GTP::setup_default_parameters();
    cfg_num_threads = std::min(cfg_num_threads, MAX_CPUS);
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto value = [&](const char* what) -> const char* {
            if (i + 1 >= argc) { std::cout << "Missing value for " << what << std::endl; usage(); exit(EXIT_FAILURE); }
            return argv[++i];
        };
        if (a == "-h" || a == "--help") { usage(); return EXIT_SUCCESS; }
        else if (a == "-g" || a == "--gtp") gtp_mode = true;
        else if (a == "-q" || a == "--quiet") cfg_quiet = true;
        else if (a == "-t" || a == "--threads") {
            const int t = atoi(value("--threads"));
            if (t > MAX_CPUS) myprintf("Clamping threads to maximum = %d\n", MAX_CPUS);
            cfg_num_threads = std::max(1, std::min(t, MAX_CPUS));
            myprintf("Using %d thread(s).\n", cfg_num_threads);
        } else if (a == "-p" || a == "--playouts") { cfg_max_playouts = atoi(value("--playouts")); playouts_set = true; }
        else if (a == "-b" || a == "--lagbuffer") {
            const int lag = atoi(value("--lagbuffer"));
            if (lag != cfg_lagbuffer_cs) { myprintf("Using per-move time margin of %.2fs.\n", lag / 100.0f); cfg_lagbuffer_cs = lag; }
        } else if (a == "-l" || a == "--logfile") {
            cfg_logfile = value("--logfile");
            myprintf("Logging to %s.\n", cfg_logfile.c_str());
            cfg_logfile_handle = fopen(cfg_logfile.c_str(), "a");
        } else if (a == "-k" || a == "--komiadjust") { myprintf("Adjusting komi for territory scoring rules.\n"); cfg_komi_adjust = true; }
        else if (a == "--noponder") { cfg_allow_pondering = false; noponder = true; }
        else if (a == "--nonets") cfg_enable_nets = false;
        // how often the search consults the nets (the reference exposes these under USE_TUNER,
        // Leela.cpp:57-75; its GPU defaults are GTP.cpp:70-73). Lower = more evaluations per playout.
        else if (a == "--mature_threshold") cfg_mature_threshold = atoi(value("--mature_threshold"));
        else if (a == "--eval_thresh") cfg_eval_thresh = atoi(value("--eval_thresh"));
        else if (a == "--extra_symmetry") cfg_extra_symmetry = atoi(value("--extra_symmetry"));
        else if (a == "--nobook") cfg_allow_book = false;
#ifndef LB2_REFERENCE_BUILD
        else if (a == "--gpu") cfg_gpus.push_back(atoi(value("--gpu")));
        else if (a == "--weights") leela_b200::set_weights_path(value("--weights"));
        else if (a == "--own-planes") leela_b200::set_planes_mode(1);
        else if (a == "--ref-planes") leela_b200::set_planes_mode(0);
        else if (a == "--check-planes") leela_b200::set_planes_mode(2);
        else if (a == "--noselftest") no_selftest = true;
        else if (a == "--max-outstanding") leela_b200::set_max_outstanding(atoi(value("--max-outstanding")));
        else if (a == "--batch") batch = atol(value("--batch"));
        else if (a == "--precise") precise = true;
        else if (a == "--queue-linger") linger = atol(value("--queue-linger"));
        else if (a == "--dump-planes") {
            dump_out = value("--dump-planes");
            dump_n = atoi(value("--dump-planes N"));
            dump_seed = (uint32_t)strtoul(value("--dump-planes SEED"), nullptr, 10);
        }
#endif
        else { std::cout << "Unrecognized argument: " << a << std::endl; usage(); return EXIT_FAILURE; }
    }
    if (playouts_set && !noponder) {
        myprintf("Nonsensical options: Playouts are restricted but thinking on the opponent's time is still allowed. "
                 "Add --noponder if you want a weakened engine.\n");
        return EXIT_FAILURE;
    }

    std::cout.setf(std::ios::unitbuf);
    std::cerr.setf(std::ios::unitbuf);
    std::cin.setf(std::ios::unitbuf);
    setbuf(stdout, NULL);
    setbuf(stderr, NULL);
    setbuf(stdin, NULL);

    thread_pool.initialize(cfg_num_threads);
    std::unique_ptr<Random> rng(new Random(5489));   // deterministic hashing, Leela.cpp:279-280
    Zobrist::init_zobrist(*rng);
    AttribScores::get_attribscores();
    Matcher::get_Matcher();

#ifndef LB2_REFERENCE_BUILD
    if (dump_out) return dump_planes(dump_out, dump_n, dump_seed);
#endif
    if (cfg_enable_nets) {
        Network::get_Network();
#ifndef LB2_REFERENCE_BUILD
        if ((batch > 0 && lb2_set_option(leela_b200::context(), "max_batch", batch)) ||
            (linger >= 0 && lb2_set_option(leela_b200::context(), "queue_linger", linger)) ||
            (precise && lb2_set_option(leela_b200::context(), "precise", 1))) {
            myprintf("%s\n", lb2_last_error());
            return EXIT_FAILURE;
        }
#endif
        myprintf("Network backend: %s\n", Network::get_Network()->get_backend().c_str());
#ifndef LB2_REFERENCE_BUILD
        atexit(print_evaluator_stats);
#endif
    }

    std::unique_ptr<GameState> maingame(new GameState);
    maingame->init_game(19, 7.5f);
#ifndef LB2_REFERENCE_BUILD
    if (cfg_enable_nets && !no_selftest && !leela_b200::self_test(*maingame)) return EXIT_FAILURE;
#endif

    std::string input;
    for (;;) {
        if (!gtp_mode) {
            maingame->display_state();
            std::cout << "Leela: ";
        }
        if (!std::getline(std::cin, input)) break;
        Utils::log_input(input);
        GTP::execute(*maingame, input);
        if (cfg_logfile_handle) {   // force a flush of the logfile
            fclose(cfg_logfile_handle);
            cfg_logfile_handle = fopen(cfg_logfile.c_str(), "a");
        }
    }
    return 0;
}
