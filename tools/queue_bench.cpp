// Single-position request load on the submit queue of libleela_b200.so — what the search's threads do through
// Network::get_value / async_scored_moves, without the feature gathering in front of it:
//   T threads, each with ONE request outstanding: lb2_submit_value (or _policy) of 1 position, wait for the callback, repeat.
// Reports requests/s and the mean device batch. Weights: the synthetic nets of the engine's weight file.
//   build: g++ -O2 -std=c++17 -Iinclude -o tools/_variants/queue_bench tools/queue_bench.cpp -Lleela_b200 -lleela_b200 -lpthread -Wl,-rpath,$PWD/leela_b200
//   run:   tools/_variants/queue_bench <weights.lb2w> [threads=64] [seconds=3] [gpus=1] [policy_every=6] [queue_linger=1]
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../include/leela_b200.h"

namespace {

// the engine's weight file (leela_b200/fileio.py:write_weights): "LB2WGT01", then per net: kind, n_conv, n_ip, then per
// conv {k, c_in, c_out, weights, biases}, per ip {n_in, n_out, weights, biases}; all int32 / float32 little endian
bool load_weights(lb2_ctx* ctx, const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) { perror(path); return false; }
    char magic[8];
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "LB2WGT01", 8)) { fprintf(stderr, "%s: not a weights file\n", path); return false; }
    int32_t n_nets = 0;
    if (fread(&n_nets, 4, 1, f) != 1) return false;
    for (int i = 0; i < n_nets; i++) {
        int32_t hdr[3];
        if (fread(hdr, 4, 3, f) != 3) return false;
        lb2_net* net = nullptr;
        if (lb2_net_create(ctx, hdr[0], &net)) { fprintf(stderr, "%s\n", lb2_last_error()); return false; }
        for (int l = 0; l < hdr[1]; l++) {
            int32_t g[3];
            if (fread(g, 4, 3, f) != 3) return false;
            std::vector<float> w((size_t)g[0] * g[0] * g[1] * g[2]), b(g[2]);
            if (fread(w.data(), 4, w.size(), f) != w.size() || fread(b.data(), 4, b.size(), f) != b.size()) return false;
            if (lb2_net_push_conv(net, g[0], g[1], g[2], w.data(), b.data())) { fprintf(stderr, "%s\n", lb2_last_error()); return false; }
        }
        for (int l = 0; l < hdr[2]; l++) {
            int32_t g[2];
            if (fread(g, 4, 2, f) != 2) return false;
            std::vector<float> w((size_t)g[0] * g[1]), b(g[1]);
            if (fread(w.data(), 4, w.size(), f) != w.size() || fread(b.data(), 4, b.size(), f) != b.size()) return false;
            if (lb2_net_push_ip(net, g[0], g[1], w.data(), b.data())) { fprintf(stderr, "%s\n", lb2_last_error()); return false; }
        }
        if (lb2_net_finalize(net)) { fprintf(stderr, "%s\n", lb2_last_error()); return false; }
    }
    fclose(f);
    return true;
}

struct Waiter {
    std::mutex mu;
    std::condition_variable cv;
    bool done = false;
    int status = 0;
    static void signal(void* user, int status) {
        Waiter* w = static_cast<Waiter*>(user);
        std::lock_guard<std::mutex> lk(w->mu);
        w->status = status;
        w->done = true;
        w->cv.notify_one();
    }
    int wait() {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return done; });
        done = false;
        return status;
    }
};

}  // namespace

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s weights.lb2w [threads] [seconds] [gpus] [policy_every]\n", argv[0]); return 2; }
    const int threads = argc > 2 ? atoi(argv[2]) : 64;
    const double seconds = argc > 3 ? atof(argv[3]) : 3.0;
    const int gpus = argc > 4 ? atoi(argv[4]) : 1;
    const int policy_every = argc > 5 ? atoi(argv[5]) : 6;   // one policy request per this many requests (netbench: 2000 : 10000)
    const int linger = argc > 6 ? atoi(argv[6]) : 1;
    std::vector<int> ids;
    for (int i = 0; i < gpus; i++) ids.push_back(i);
    lb2_ctx* ctx = nullptr;
    if (lb2_init(ids.data(), gpus, &ctx)) { fprintf(stderr, "lb2_init: %s\n", lb2_last_error()); return 1; }
    if (!load_weights(ctx, argv[1])) return 1;
    lb2_set_option(ctx, "queue_linger", linger);
    // any bit patterns are valid planes; a few distinct ones so that results differ
    std::vector<uint32_t> planes(16 * 361);
    uint64_t z = 88172645463325252ull;
    for (auto& p : planes) { z ^= z << 13; z ^= z >> 7; z ^= z << 17; p = (uint32_t)z & (uint32_t)(z >> 32); }
    {   // warm-up: workspaces, graphs
        std::vector<float> probs(16 * 361), win(16);
        std::vector<uint8_t> rot(16, 0);
        for (int i = 0; i < 3; i++)
            if (lb2_eval_both(ctx, planes.data(), planes.data(), rot.data(), 16, 0.75f, probs.data(), win.data())) { fprintf(stderr, "%s\n", lb2_last_error()); return 1; }
    }
    std::atomic<bool> stop{false};
    std::atomic<long> total{0}, failed{0};
    std::vector<std::thread> pool;
    const long pos0 = lb2_get_option(ctx, "stat_positions"), bat0 = lb2_get_option(ctx, "stat_batches");
    const auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&, t]() {
            Waiter w;
            float probs[361], win = 0;
            long mine = 0;
            for (long i = t; !stop.load(std::memory_order_relaxed); i++) {
                const uint32_t* p = planes.data() + (i % 16) * 361;
                const uint8_t rot = (uint8_t)(i & 7);
                const int rc = (policy_every > 0 && i % policy_every == 0) ? lb2_submit_policy(ctx, p, &rot, 1, 0.75f, probs, Waiter::signal, &w)
                                                                           : lb2_submit_value(ctx, p, &rot, 1, &win, Waiter::signal, &w);
                if (rc || w.wait()) { failed++; break; }
                mine++;
            }
            total += mine;
        });
    std::this_thread::sleep_for(std::chrono::duration<double>(seconds));
    stop = true;
    for (auto& th : pool) th.join();
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    lb2_drain(ctx);
    const long pos = lb2_get_option(ctx, "stat_positions") - pos0, bat = lb2_get_option(ctx, "stat_batches") - bat0;
    printf("{\"threads\": %d, \"gpus\": %d, \"seconds\": %.2f, \"requests\": %ld, \"requests_per_s\": %.0f, \"device_batches\": %ld, "
           "\"mean_device_batch\": %.1f, \"failed\": %ld, \"policy_every\": %d, \"queue_linger\": %d}\n",
           threads, gpus, dt, total.load(), total.load() / dt, bat, bat ? (double)pos / bat : 0.0, failed.load(), policy_every, linger);
    lb2_destroy(ctx);
    return failed ? 1 : 0;
}
