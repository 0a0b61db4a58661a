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
// liberty counts, so a board that re-derives strings, sizes and liberties by flood fill gives the same
// answers — the ladder reader's decisions depend only on liberty counts and on sets of points, never on
// the traversal order of a string. Round 1 re-analysed the whole board after every stone (0.12 ms per
// position, two thirds of it in the ladder readers and in the 2 x ~250 liberties-after-a-move fills);
// now a move re-fills only the strings it touches (the one it joins, the ones it captures, their
// neighbours), the liberties-after-a-move count for the planes stops at the 6 the planes distinguish,
// and visited marks are generation stamps instead of cleared arrays.
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

constexpr int kLibList = 6;        // liberties remembered per string (the planes distinguish 1..5 and ">= 6")
constexpr int kMaxGroups = 1536;   // string ids are never recycled: <= 361 at the start + one per stone played in a ladder read

// visited marks of the flood fills: a point is marked when mark[p] == stamp (no clearing between fills)
struct Scratch {
    uint16_t mark[SQ];
    uint16_t stamp = 0;
    int16_t stack[N * N + 4];
    uint16_t next() {
        if (++stamp == 0) { memset(mark, 0, sizeof mark); stamp = 1; }
        return stamp;
    }
};
thread_local Scratch g_scratch;

struct Board {
    uint8_t sq[SQ];
    int16_t group[SQ];             // string id per stone, -1 otherwise
    int16_t libs[kMaxGroups];      // per string: number of distinct empty points adjacent to it
    int16_t stones[kMaxGroups];    // per string: number of stones
    int16_t lib_pts[kMaxGroups][kLibList];   // per string: its liberties, all of them when there are at most kLibList
    int16_t n_groups;
    uint8_t tomove;

    Board() = default;
    Board(const Board& o) { *this = o; }
    Board& operator=(const Board& o) {   // (the ladder readers copy the board per branch: only the string ids in use)
        memcpy(sq, o.sq, sizeof sq);
        memcpy(group, o.group, sizeof group);
        memcpy(libs, o.libs, (size_t)o.n_groups * sizeof libs[0]);
        memcpy(stones, o.stones, (size_t)o.n_groups * sizeof stones[0]);
        memcpy(lib_pts, o.lib_pts, (size_t)o.n_groups * sizeof lib_pts[0]);
        n_groups = o.n_groups;
        tomove = o.tomove;
        return *this;
    }

    void clear() {
        for (int v = 0; v < SQ; v++) sq[v] = BORDER;
        for (int i = 0; i < N * N; i++) sq[vertex_of(i)] = EMPTY;
        n_groups = 0;
    }
    // the string containing the stone at v0 gets id g: its stones, size and true liberties by flood fill
    void fill_group(int v0, int g) {
        Scratch& S = g_scratch;
        const uint16_t st = S.next();
        const uint8_t c = sq[v0];
        int sp = 0, n_st = 0, n_lib = 0;
        S.stack[sp++] = (int16_t)v0;
        S.mark[v0] = st;
        group[v0] = (int16_t)g;
        while (sp) {
            const int v = S.stack[--sp];
            n_st++;
            for (int k = 0; k < 4; k++) {
                const int a = v + kDirs[k];
                if (S.mark[a] == st) continue;
                if (sq[a] == EMPTY) { S.mark[a] = st; if (n_lib < kLibList) lib_pts[g][n_lib] = (int16_t)a; n_lib++; }
                else if (sq[a] == c) { S.mark[a] = st; group[a] = (int16_t)g; S.stack[sp++] = (int16_t)a; }
            }
        }
        libs[g] = (int16_t)n_lib;
        stones[g] = (int16_t)n_st;
    }
    int new_group() { return n_groups < kMaxGroups - 1 ? n_groups++ : kMaxGroups - 1; }   // (the bound is never reached on a 19 x 19 board)

    // strings, their sizes and true liberties of the whole board
    void analyze() {
        for (int v = 0; v < SQ; v++) group[v] = -1;
        n_groups = 0;
        for (int i = 0; i < N * N; i++) {
            const int v0 = vertex_of(i);
            if (sq[v0] <= WHITE && group[v0] < 0) fill_group(v0, new_group());
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

    // take the string containing v0 off the board; the strings of the other colour around it get their liberties back
    void remove_group(int v0) {
        Scratch& S = g_scratch;
        const uint8_t c = sq[v0];
        const int g = group[v0];
        int16_t members[N * N];
        int n = 0, sp = 0;
        S.stack[sp++] = (int16_t)v0;
        sq[v0] = EMPTY;                 // (emptied as it is visited: doubles as the visited mark)
        while (sp) {
            const int v = S.stack[--sp];
            members[n++] = (int16_t)v;
            group[v] = -1;
            for (int k = 0; k < 4; k++) {
                const int a = v + kDirs[k];
                if (sq[a] == c && group[a] == g) { sq[a] = EMPTY; S.stack[sp++] = (int16_t)a; }
            }
        }
        // every enemy string that touched a removed stone: re-fill it once (its liberties changed)
        int16_t redone[N * N];
        int n_redone = 0;
        for (int i = 0; i < n; i++)
            for (int k = 0; k < 4; k++) {
                const int a = members[i] + kDirs[k];
                if (sq[a] == (uint8_t)!c) {
                    const int ga = group[a];
                    bool dup = false;
                    for (int j = 0; j < n_redone && !dup; j++) dup = redone[j] == ga;
                    if (!dup) { redone[n_redone++] = (int16_t)ga; fill_group(a, ga); }
                }
            }
    }

    // update_board_fast (FastBoard.cpp:812-874): place a stone, capture, detect multi-stone
    // suicide; ko is not a concept here. A play into an opponent eye (all four neighbours opponent
    // or border) takes the capture-only path of update_board_eye. Only the strings the move touches
    // are re-derived.
    void play(int c, int v) {
        const bool eyeplay = colour_neighbours(!c, v) == 4;
        sq[v] = (uint8_t)c;
        // the enemy strings around v each lose the liberty v (re-filled, so that their liberty lists stay exact)
        int enemy[4], n_enemy = 0;
        for (int k = 0; k < 4; k++) {
            const int a = v + kDirs[k];
            if (sq[a] == (uint8_t)!c) {
                bool dup = false;
                for (int j = 0; j < n_enemy; j++) dup |= enemy[j] == group[a];
                if (!dup) { enemy[n_enemy++] = group[a]; fill_group(a, group[a]); }
            }
        }
        fill_group(v, new_group());     // the stone and the friendly strings it joins
        for (int k = 0; k < 4; k++) {
            const int a = v + kDirs[k];
            if (sq[a] == (uint8_t)!c && libs[group[a]] == 0) remove_group(a);   // (re-fills the string of v when it borders the capture)
        }
        if (!eyeplay && libs[group[v]] == 0) remove_group(v);
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
        int found = 0;
        for_each_stone(v, [&](int p) {
            for (int k = 0; k < 4 && !found; k++)
                if (sq[p + kDirs[k]] == EMPTY) found = p + kDirs[k];
            return found == 0;
        });
        return found;
    }

    // visits the stones of the string containing v0 until f returns false
    template <class F>
    void for_each_stone(int v0, F f) const {
        Scratch& S = g_scratch;
        const uint16_t st = S.next();
        const uint8_t c = sq[v0];
        int sp = 0;
        S.stack[sp++] = (int16_t)v0;
        S.mark[v0] = st;
        while (sp) {
            const int p = S.stack[--sp];
            if (!f(p)) return;
            for (int k = 0; k < 4; k++) {
                const int a = p + kDirs[k];
                if (sq[a] == c && S.mark[a] != st) { S.mark[a] = st; S.stack[sp++] = (int16_t)a; }
            }
        }
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
                for_each_stone(a, [&](int p) {
                    for (int kk = 0; kk < 4; kk++)
                        if (sq[p + kDirs[kk]] == EMPTY) add(p + kDirs[kk]);
                    return n <= 2;
                });
            }
            if (n > 2) return false;
        }
        return true;
    }

    // after_liberties_color (FastBoard.cpp:2511-2522): liberties of the string that playing v creates,
    // 0 for a suicide. The reference copies the board and plays the move; the same number comes out of
    // one local flood fill from v in which the enemy neighbour strings the move captures (those whose
    // only liberty is v) already count as empty points.
    // `cap`: stop counting there (the planes only distinguish 1..5 and ">= 6"; the ladder reader needs the exact number)
    int after_liberties(int c, int v, int cap = 1 << 20) const {
        if (is_suicide(v, c)) return 0;
        int captured[4], n_cap = 0, n_friends = 0, n_empty = 0;
        for (int k = 0; k < 4; k++) {
            const int a = v + kDirs[k];
            if (sq[a] == EMPTY) n_empty++;
            else if (sq[a] == (uint8_t)!c) { if (libs[group[a]] == 1) captured[n_cap++] = group[a]; }
            else if (sq[a] == c) {
                if (libs[group[a]] > cap) return cap;   // joining a string with cap + 1 liberties (one of them is v) settles it at once
                n_friends++;
            }
        }
        if (!n_friends && !n_cap) return std::min(n_empty, cap);   // a lone stone that captures nothing: its empty neighbours
        if (!n_cap && cap <= kLibList) {
            // nothing captured and every string joined has at most `cap` liberties, all of them listed: the answer is the
            // size of the union of the empty neighbours of v and those lists, without v
            int pts[4 + 4 * kLibList], n = 0;
            auto add = [&](int q) {
                if (q == v) return;
                for (int i = 0; i < n; i++) if (pts[i] == q) return;
                pts[n++] = q;
            };
            int seen_g[4], n_seen = 0;
            for (int k = 0; k < 4; k++) {
                const int a = v + kDirs[k];
                if (sq[a] == EMPTY) add(a);
                else if (sq[a] == c) {
                    const int g = group[a];
                    bool dup = false;
                    for (int i = 0; i < n_seen; i++) dup |= seen_g[i] == g;
                    if (dup) continue;
                    seen_g[n_seen++] = g;
                    for (int i = 0; i < libs[g]; i++) add(lib_pts[g][i]);
                }
            }
            return std::min(n, cap);
        }
        auto vacated = [&](int p) {
            for (int i = 0; i < n_cap; i++) if (group[p] == captured[i]) return true;
            return false;
        };
        Scratch& S = g_scratch;
        const uint16_t st = S.next();
        int sp = 0, n_lib = 0;
        S.stack[sp++] = (int16_t)v;
        S.mark[v] = st;
        while (sp) {
            const int p = S.stack[--sp];
            for (int k = 0; k < 4; k++) {
                const int a = p + kDirs[k];
                if (S.mark[a] == st) continue;
                if (sq[a] == c) { S.mark[a] = st; S.stack[sp++] = (int16_t)a; }
                else if (sq[a] == EMPTY || (sq[a] == (uint8_t)!c && n_cap && vacated(a))) {
                    S.mark[a] = st;
                    if (++n_lib >= cap) return cap;
                }
            }
        }
        return n_lib;
    }

    // The planes ask for the liberties-after-a-move of BOTH colours at every empty point: one scan of the four neighbours serves
    // both. Same answers as after_liberties(c, v, cap) for c = BLACK, WHITE (cap <= kLibList).
    void after_liberties_both(int v, int cap, int out[2]) const {
        int nb_sq[4], nb_g[4], nb_l[4], n_empty = 0;
        for (int k = 0; k < 4; k++) {
            const int a = v + kDirs[k];
            nb_sq[k] = sq[a];
            nb_g[k] = sq[a] <= WHITE ? group[a] : -1;
            nb_l[k] = nb_g[k] >= 0 ? libs[nb_g[k]] : kFar;
            n_empty += sq[a] == EMPTY;
        }
        for (int c = 0; c < 2; c++) {
            int n_friends = 0, n_cap = 0;
            bool big = false, live_friend = false;
            for (int k = 0; k < 4; k++) {
                if (nb_sq[k] == c) { n_friends++; big |= nb_l[k] > cap; live_friend |= nb_l[k] > 1; }
                else if (nb_sq[k] == (c ^ 1)) n_cap += nb_l[k] == 1;
            }
            if (!n_empty && !live_friend && !n_cap) { out[c] = 0; continue; }      // is_suicide
            if (big) { out[c] = cap; continue; }
            if (!n_friends && !n_cap) { out[c] = std::min(n_empty, cap); continue; }
            if (n_cap) { out[c] = after_liberties(c, v, cap); continue; }           // captures: the flood fill
            int pts[4 + 4 * kLibList], n = 0;
            auto add = [&](int q) {
                if (q == v) return;
                for (int i = 0; i < n; i++) if (pts[i] == q) return;
                pts[n++] = q;
            };
            int seen_g[4], n_seen = 0;
            for (int k = 0; k < 4; k++) {
                if (nb_sq[k] == EMPTY) add(v + kDirs[k]);
                else if (nb_sq[k] == c) {
                    const int g = nb_g[k];
                    bool dup = false;
                    for (int i = 0; i < n_seen; i++) dup |= seen_g[i] == g;
                    if (dup) continue;
                    seen_g[n_seen++] = g;
                    for (int i = 0; i < nb_l[k]; i++) add(lib_pts[g][i]);
                }
            }
            out[c] = std::min(n, cap);
        }
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

    // can_kill_neighbours (FastBoard.cpp:2587-2613): some enemy string touching the string of v0 is in atari
    bool can_kill_neighbours(int v0) const {
        const uint8_t enemy = (uint8_t)!sq[v0];
        bool found = false;
        for_each_stone(v0, [&](int p) {
            for (int k = 0; k < 4; k++) {
                const int a = p + kDirs[k];
                if (sq[a] == enemy && libs[group[a]] <= 1) found = true;
            }
            return !found;
        });
        return found;
    }

    // check_losing_ladder (FastBoard.cpp:2647-2837): `c` (== tomove) extends at v out of atari;
    // does the attacker capture anyway by chasing?
    bool losing_ladder(int c, int v, int branching = 0) const {
        if (branching > 5) return false;
        const int elib = minimum_enemy_libs(c, v);
        if (elib == 0 || elib == 1) return false;                 // the move captures something
        // the friendly strings in atari next to v: more than one means we are connecting, not running
        int crit[4], crit_at = 0, n_crit = 0;
        for (int k = 0; k < 4; k++) {
            const int a = v + kDirs[k];
            if (sq[a] == c && libs_at(a) <= 1) {
                bool dup = false;
                for (int i = 0; i < n_crit; i++) dup |= crit[i] == group[a];
                if (!dup) { crit[n_crit++] = group[a]; crit_at = a; }
            }
        }
        if (n_crit != 1) return false;    // (the reference asserts n_crit > 0; callers guarantee it)
        if (can_kill_neighbours(crit_at)) return false;           // an atari-giving stone can be captured instead

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
            t.for_each_stone(atari, [&](int p) {
                for (int k = 0; k < 4 && nl < 2; k++) {
                    const int a = p + kDirs[k];
                    if (t.sq[a] == EMPTY && (nl == 0 || lib[0] != a)) lib[nl++] = a;
                }
                return nl < 2;
            });
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
        bool captures = false;   // does a stone at v capture anything?
        for (int k = 0; k < 4; k++) {
            const int a = v + kDirs[k];
            captures |= sq[a] == (uint8_t)!c && libs[group[a]] == 1;
        }
        for (int k = 0; k < 4; k++) {
            const int a = v + kDirs[k];
            if (sq[a] != (uint8_t)!c || libs[group[a]] != 2) continue;
            if (self_atari(c, v)) continue;
            if (!captures) {
                // the board after the move differs only by the stone at v: the string keeps its other liberty, and the ladder is
                // only read when that escape point has exactly two empty neighbours — decided here, before any board is copied
                const int g = group[a];
                const int escape = lib_pts[g][0] == v ? lib_pts[g][1] : lib_pts[g][0];
                int en = empty_neighbours(escape);
                for (int kk = 0; kk < 4; kk++) en -= (escape + kDirs[kk] == v);
                if (en != 2) continue;
            }
            Board t = *this;
            t.play(t.tomove, v);
            const int escape = t.in_atari(a);
            if (!escape) continue;   // (the capture gave the string room: the reference reads "0 empty neighbours" off its border vertex 0)
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
            int after[2];
            b.after_liberties_both(v, 6, after);
            p.after_own = (int16_t)after[c];
            p.after_opp = (int16_t)after[!c];
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
