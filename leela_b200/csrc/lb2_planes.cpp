// lb2_planes.cpp — feature planes of the policy and value nets from a RAW position (stones, side
// to move, ko point, last two moves, komi), with a Go board of our own: no reference code is needed
// to produce the evaluator's inputs.
//
// Reproduces, bit for bit, what the reference computes with
//   Network::gather_features_policy / _value          Network.cpp:883-1201
// on top of its FastBoard queries
//   count_rliberties / count_pliberties               FastBoard.cpp:2482-2509, 242-251
//   after_liberties(_color), is_suicide               FastBoard.cpp:2511-2528, 191-240
//   saving_size, self_atari, kill_or_connect          FastBoard.cpp:1273-1295, 1401-1471, 1342-1355
//   check_losing_ladder, check_winning_ladder         FastBoard.cpp:2647-2837, 2530-2564
//   minimum_elib_count, critical_neighbours, can_kill_neighbours, in_atari, update_board_fast
// The reference keeps strings incrementally (parent / next / libs arrays); its m_libs are TRUE
// liberty counts, so a board that simply re-derives strings, sizes and liberties by flood fill after
// every change gives the same answers — the ladder reader's decisions depend only on liberty counts
// and on sets of points, never on the traversal order of a string. Positions are tiny (361 points),
// a full re-analysis costs about a microsecond.
//
// Checked against the reference's planes on thousands of seeded self-play positions
// (tests/test_planes.py); exported through the C ABI as lb2_planes_from_position.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>

#include "../../include/leela_b200.h"

namespace {

constexpr int N = 19, W = N + 2, SQ = W * W;          // 21 x 21 with a border ring
enum : uint8_t { BLACK = 0, WHITE = 1, EMPTY = 2, BORDER = 3 };
constexpr int kDirs[4] = {-W, +1, +W, -1};
constexpr int kFar = 16384;                            // "liberties" of empty / border squares (FastBoard.cpp:186-188)

inline int vertex_of(int idx) { return (idx / N + 1) * W + (idx % N + 1); }   // FastBoard::get_vertex
inline int idx_of(int v) { return (v / W - 1) * N + (v % W - 1); }

struct Board {
    uint8_t sq[SQ];
    int16_t group[SQ];      // string id per stone, -1 otherwise
    int16_t libs[SQ / 2];   // per string: number of distinct empty points adjacent to it
    int16_t stones[SQ / 2]; // per string: number of stones
    uint8_t tomove;

    void clear() {
        for (int v = 0; v < SQ; v++) sq[v] = BORDER;
        for (int i = 0; i < N * N; i++) sq[vertex_of(i)] = EMPTY;
    }

    // strings, their sizes and true liberties
    void analyze() {
        int16_t stack[N * N];
        int16_t seen_lib[SQ];   // last string that counted this empty point
        memset(seen_lib, -1, sizeof seen_lib);
        for (int v = 0; v < SQ; v++) group[v] = -1;
        int n_groups = 0;
        for (int i = 0; i < N * N; i++) {
            const int v0 = vertex_of(i);
            if (sq[v0] > WHITE || group[v0] >= 0) continue;
            const int g = n_groups++;
            const uint8_t c = sq[v0];
            int sp = 0, n_st = 0, n_lib = 0;
            stack[sp++] = (int16_t)v0;
            group[v0] = (int16_t)g;
            while (sp) {
                const int v = stack[--sp];
                n_st++;
                for (int k = 0; k < 4; k++) {
                    const int a = v + kDirs[k];
                    if (sq[a] == EMPTY) {
                        if (seen_lib[a] != g) { seen_lib[a] = (int16_t)g; n_lib++; }
                    } else if (sq[a] == c && group[a] < 0) {
                        group[a] = (int16_t)g;
                        stack[sp++] = (int16_t)a;
                    }
                }
            }
            libs[g] = (int16_t)n_lib;
            stones[g] = (int16_t)n_st;
        }
    }

    int libs_at(int v) const { return group[v] >= 0 ? libs[group[v]] : kFar; }   // m_libs[m_parent[v]]
    int empty_neighbours(int v) const {                                            // count_pliberties
        int n = 0;
        for (int k = 0; k < 4; k++) n += sq[v + kDirs[k]] == EMPTY;
        return n;
    }
    int colour_neighbours(int c, int v) const {   // count_neighbours: the border counts as both colours
        int n = 0;
        for (int k = 0; k < 4; k++) { const uint8_t s = sq[v + kDirs[k]]; n += (s == c || s == BORDER); }
        return n;
    }

    void remove_group(int g) {
        for (int v = 0; v < SQ; v++)
            if (group[v] == g) sq[v] = EMPTY;
    }

    // update_board_fast (FastBoard.cpp:812-874): place a stone, capture, detect multi-stone
    // suicide; ko is not a concept here. A play into an opponent eye (all four neighbours opponent
    // or border) takes the capture-only path of update_board_eye.
    void play(int c, int v) {
        const bool eyeplay = colour_neighbours(!c, v) == 4;
        sq[v] = (uint8_t)c;
        analyze();
        for (int k = 0; k < 4; k++) {
            const int a = v + kDirs[k];
            if (sq[a] == (uint8_t)!c && libs[group[a]] == 0) { remove_group(group[a]); analyze(); }
        }
        if (!eyeplay && libs[group[v]] == 0) { remove_group(group[v]); analyze(); }
    }

    // is_suicide (FastBoard.cpp:191-240). The reference's early exits already decide everything:
    // a point with an empty neighbour, next to a friendly string with a spare liberty, or next to an
    // enemy string in atari is playable; otherwise the stone (and every friendly string it joins,
    // each of which has this point as its only liberty) ends up without liberties and captures nothing.
    bool is_suicide(int v, int c) const {
        if (empty_neighbours(v)) return false;
        for (int k = 0; k < 4; k++) {
            const int a = v + kDirs[k];
            if (sq[a] == c) {
                if (libs_at(a) > 1) return false;      // connecting to a live group
            } else if (libs_at(a) <= 1) {
                return false;                          // killing a neighbour
            }
        }
        return true;
    }

    // in_atari: the single liberty of the string at v, 0 if it has more than one (FastBoard.cpp:1179-1206)
    int in_atari(int v) const {
        const int g = group[v];
        if (libs[g] > 1) return 0;
        for (int p = 0; p < SQ; p++)
            if (group[p] == g)
                for (int k = 0; k < 4; k++)
                    if (sq[p + kDirs[k]] == EMPTY) return p + kDirs[k];
        return 0;
    }

    // kill_or_connect (FastBoard.cpp:1342-1355)
    bool kill_or_connect(int c, int v) const {
        for (int k = 0; k < 4; k++) {
            const int a = v + kDirs[k];
            const int l = libs_at(a);
            if ((l <= 1 && sq[a] == (uint8_t)!c) || (l >= 3 && sq[a] == c)) return true;
        }
        return false;
    }

    // self_atari (FastBoard.cpp:1401-1471): does the group formed by playing v end up with at most
    // one liberty? The point itself, its empty neighbours and the liberties of friendly neighbour
    // strings (those with more than the one liberty that is v) are collected; three or more = safe.
    bool self_atari(int c, int v) const {
        if (empty_neighbours(v) >= 2) return false;
        if (kill_or_connect(c, v)) return false;
        if (colour_neighbours(c, v) == 0) return true;   // (border squares count as neighbours of both colours)
        int pts[8], n = 0;
        pts[n++] = v;
        auto add = [&](int p) {
            for (int i = 0; i < n; i++) if (pts[i] == p) return;
            if (n < 8) pts[n++] = p;
        };
        for (int k = 0; k < 4; k++) {
            const int a = v + kDirs[k];
            if (sq[a] == EMPTY) {
                add(a);
            } else if (sq[a] == c && libs_at(a) > 1) {
                const int g = group[a];
                for (int p = 0; p < SQ && n <= 2; p++)
                    if (group[p] == g)
                        for (int kk = 0; kk < 4; kk++)
                            if (sq[p + kDirs[kk]] == EMPTY) add(p + kDirs[kk]);
            }
            if (n > 2) return false;
        }
        return true;
    }

    // after_liberties_color (FastBoard.cpp:2511-2522): liberties of the string that playing v creates,
    // 0 for a suicide. The reference copies the board and plays the move; the same number comes out of
    // one local flood fill from v in which the enemy neighbour strings the move captures (those whose
    // only liberty is v) already count as empty points.
    int after_liberties(int c, int v) const {
        if (is_suicide(v, c)) return 0;
        int captured[4], n_cap = 0;
        for (int k = 0; k < 4; k++) {
            const int a = v + kDirs[k];
            if (sq[a] == (uint8_t)!c && libs[group[a]] == 1) captured[n_cap++] = group[a];
        }
        auto vacated = [&](int p) {
            for (int i = 0; i < n_cap; i++) if (group[p] == captured[i]) return true;
            return false;
        };
        uint8_t seen[SQ];
        memset(seen, 0, sizeof seen);
        int16_t stack[N * N];
        int sp = 0, n_lib = 0;
        stack[sp++] = (int16_t)v;
        seen[v] = 1;
        while (sp) {
            const int p = stack[--sp];
            for (int k = 0; k < 4; k++) {
                const int a = p + kDirs[k];
                if (seen[a]) continue;
                if (sq[a] == c) { seen[a] = 1; stack[sp++] = (int16_t)a; }
                else if (sq[a] == EMPTY || (sq[a] == (uint8_t)!c && n_cap && vacated(a))) { seen[a] = 1; n_lib++; }
            }
        }
        return n_lib;
    }

    // minimum_elib_count (FastBoard.cpp:2429-2443)
    int minimum_enemy_libs(int c, int v) const {
        int m = 100;
        for (int k = 0; k < 4; k++) {
            const int a = v + kDirs[k];
            if (sq[a] == (uint8_t)!c) m = std::min(m, (int)libs[group[a]]);
        }
        return m;
    }

    // saving_size > 0 (FastBoard.cpp:1273-1295): v rescues a friendly neighbour string in atari
    // without being a self-atari itself
    bool saves_something(int c, int v) const {
        for (int k = 0; k < 4; k++) {
            const int a = v + kDirs[k];
            if (sq[a] == c && libs_at(a) <= 1) {
                const int lib = in_atari(a);
                if (!self_atari(c, lib)) return true;
            }
        }
        return false;
    }

    // can_kill_neighbours (FastBoard.cpp:2587-2613): some enemy string touching string g is in atari
    bool can_kill_neighbours(int g) const {
        const uint8_t enemy = (uint8_t)!sq_of_group(g);
        for (int p = 0; p < SQ; p++)
            if (group[p] == g)
                for (int k = 0; k < 4; k++) {
                    const int a = p + kDirs[k];
                    if (sq[a] == enemy && libs[group[a]] <= 1) return true;
                }
        return false;
    }
    uint8_t sq_of_group(int g) const {
        for (int p = 0; p < SQ; p++) if (group[p] == g) return sq[p];
        return EMPTY;
    }

    // check_losing_ladder (FastBoard.cpp:2647-2837): `c` (== tomove) extends at v out of atari;
    // does the attacker capture anyway by chasing?
    bool losing_ladder(int c, int v, int branching = 0) const {
        if (branching > 5) return false;
        const int elib = minimum_enemy_libs(c, v);
        if (elib == 0 || elib == 1) return false;                 // the move captures something
        // the friendly strings in atari next to v: more than one means we are connecting, not running
        int crit[4], n_crit = 0;
        for (int k = 0; k < 4; k++) {
            const int a = v + kDirs[k];
            if (sq[a] == c && libs_at(a) <= 1) {
                bool dup = false;
                for (int i = 0; i < n_crit; i++) dup |= crit[i] == group[a];
                if (!dup) crit[n_crit++] = group[a];
            }
        }
        if (n_crit != 1) return false;    // (the reference asserts n_crit > 0; callers guarantee it)
        if (can_kill_neighbours(crit[0])) return false;           // an atari-giving stone can be captured instead

        Board t = *this;
        int atari = v;
        t.play(t.tomove, v);
        // (the side to move never changes: defender t.tomove, attacker the other. Every round adds two
        // stones, so a real chase ends within the board's 361 points; the bound only guards against
        // malformed input arriving through the ABI)
        for (int round = 0; round < 400; round++) {
            if (t.sq[atari] == EMPTY) return true;                // the extension was suicide
            const int newlibs = t.libs[t.group[atari]];
            if (newlibs == 1) return true;                        // still in atari
            if (newlibs >= 3) return false;                       // escaped
            if (t.minimum_enemy_libs(c, atari) == 1) return false;  // counter-atari on a chaser
            // the two liberties of the running string
            int lib[2], nl = 0;
            const int g = t.group[atari];
            for (int p = 0; p < SQ && nl < 2; p++)
                if (t.group[p] == g)
                    for (int k = 0; k < 4 && nl < 2; k++) {
                        const int a = p + kDirs[k];
                        if (t.sq[a] == EMPTY && (nl == 0 || lib[0] != a)) lib[nl++] = a;
                    }
            if (t.empty_neighbours(lib[0]) == 3 && t.empty_neighbours(lib[1]) == 3) return false;   // two good ways out
            // where does the attacker atari next: the liberty whose escape would gain the defender more
            int gain0 = t.after_liberties(t.tomove, lib[0]);
            int gain1 = t.after_liberties(t.tomove, lib[1]);
            const int attacker = !t.tomove;
            const bool bad0 = t.is_suicide(lib[0], attacker) || t.self_atari(attacker, lib[0]);
            const bool bad1 = t.is_suicide(lib[1], attacker) || t.self_atari(attacker, lib[1]);
            if (bad0 && bad1) return false;
            if (bad0) gain1 = gain0 + 1;
            if (bad1) gain0 = gain1 + 1;
            if (gain0 == gain1) {         // no preference: the ladder works if either atari works
                Board b0 = t;
                b0.play(attacker, lib[0]);
                if (b0.losing_ladder(c, lib[1], branching + 1)) return true;
                Board b1 = t;
                b1.play(attacker, lib[1]);
                return b1.losing_ladder(c, lib[0], branching + 1);
            }
            t.play(attacker, gain0 > gain1 ? lib[0] : lib[1]);
            atari = t.in_atari(atari);    // the defender's only move: extend again
            if (!atari) return false;     // (cannot happen on a legal board: the string was just put in atari)
            t.play(t.tomove, atari);
        }
        return false;
    }

    // check_winning_ladder (FastBoard.cpp:2530-2564): `c` (== tomove) gives atari at v on a
    // neighbouring two-liberty string whose only escape runs into a working ladder
    bool winning_ladder(int c, int v) const {
        for (int k = 0; k < 4; k++) {
            const int a = v + kDirs[k];
            if (sq[a] != (uint8_t)!c || libs[group[a]] != 2) continue;
            if (self_atari(c, v)) continue;
            Board t = *this;
            t.play(t.tomove, v);
            const int escape = t.in_atari(a);
            if (t.empty_neighbours(escape) == 2) {
                t.tomove = (uint8_t)!t.tomove;
                if (t.losing_ladder(t.tomove, escape)) return true;
            }
        }
        return false;
    }
};

// plane numbering of the two nets (Network.cpp:886-917 policy, :1050-1081 value)
struct PlaneLayout {
    int own_libs, opp_libs, libs_cap, own_after, opp_after, ladder, ladder_win, ko, last_move, prev_move, komi, line3;
};
const PlaneLayout kPolicyPlanes = {3, 8, 5, 13, 19, 25, 26, 27, 28, 29, 30, 31};
const PlaneLayout kValuePlanes = {3, 9, 6, 15, 21, 27, 28, 31, -1, -1, 29, 30};

inline uint32_t plane(int p) { return 1u << p; }
inline uint32_t count_plane(int base, int count, int cap) { return count >= 1 ? plane(base + std::min(count, cap) - 1) : 0u; }

}  // namespace

extern "C" int lb2_planes_from_position(const uint8_t* stones, int white_to_move, int ko_point, int last_move, int prev_move,
                                        float komi, uint32_t* policy_planes, uint32_t* value_planes) {
    if (!stones || (!policy_planes && !value_planes)) return LB2_ERR_INVALID;
    if (ko_point >= 361 || last_move >= 361 || prev_move >= 361) return LB2_ERR_INVALID;
    Board b;
    b.clear();
    for (int i = 0; i < 361; i++) {
        if (stones[i] > 2) return LB2_ERR_INVALID;
        if (stones[i]) b.sq[vertex_of(i)] = stones[i] == 1 ? BLACK : WHITE;
    }
    b.tomove = white_to_move ? WHITE : BLACK;
    b.analyze();
    const int c = b.tomove;
    const bool white_has_komi = std::fabs(komi) > 0.75f;

    // per point: what both nets share
    struct Point { uint8_t sq; int16_t libs, after_own, after_opp; bool ladder, ladder_win; };
    Point pt[361];
    for (int idx = 0; idx < 361; idx++) {
        const int v = vertex_of(idx);
        Point& p = pt[idx];
        p.sq = b.sq[v];
        p.libs = p.after_own = p.after_opp = 0;
        p.ladder = p.ladder_win = false;
        if (p.sq != EMPTY) {
            p.libs = b.libs[b.group[v]];
        } else {
            p.after_own = (int16_t)b.after_liberties(c, v);
            p.after_opp = (int16_t)b.after_liberties(!c, v);
            p.ladder = b.empty_neighbours(v) == 2 && b.saves_something(c, v) && b.losing_ladder(c, v);
            p.ladder_win = b.winning_ladder(c, v);
        }
    }
    for (int net = 0; net < 2; net++) {
        uint32_t* out = net == 0 ? policy_planes : value_planes;
        if (!out) continue;
        const PlaneLayout& L = net == 0 ? kPolicyPlanes : kValuePlanes;
        for (int idx = 0; idx < 361; idx++) {
            const int x = idx % 19, y = idx / 19;
            const Point& p = pt[idx];
            uint32_t bits = (x == 2 || x == 16 || y == 2 || y == 16) ? plane(L.line3) : 0u;
            if (p.sq != EMPTY) {
                const bool own = p.sq == c;
                bits |= own ? plane(1) : plane(2);
                if (p.sq == WHITE && white_has_komi) bits |= plane(L.komi);
                bits |= count_plane(own ? L.own_libs : L.opp_libs, p.libs, L.libs_cap);
            } else {
                bits |= plane(0) | count_plane(L.own_after, p.after_own, 6) | count_plane(L.opp_after, p.after_opp, 6);
                if (p.ladder) bits |= plane(L.ladder);
                if (p.ladder_win) bits |= plane(L.ladder_win);
            }
            out[idx] = bits;
        }
        if (last_move >= 0 && L.last_move >= 0) {
            out[last_move] |= plane(L.last_move);
            if (prev_move >= 0) out[prev_move] |= plane(L.prev_move);
        }
        if (ko_point >= 0) out[ko_point] |= plane(L.ko);
    }
    return LB2_OK;
}
