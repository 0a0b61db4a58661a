"""lb2_planes_from_position: feature planes from RAW positions with the library's own Go board
(leela_b200/csrc/lb2_planes.cpp) — liberties, liberties after a move, the ladder readers — against
the reference's gather_features_policy / _value + FastBoard (Network.cpp:883-1201,
FastBoard.cpp:2482-2564, 2647-2837), bit for bit, on seeded self-play positions.

The positions come out of the drop-in engine's --dump-planes (raw position + planes computed through
the reference's board queries; tests/test_engine.py pins those planes to the unmodified reference)."""
import os
import subprocess

import numpy as np
import pytest

from leela_b200 import capi, fileio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENGINE = os.path.join(ROOT, "engine", "_build", "leela_b200_engine")
needs_engine = pytest.mark.skipif(not os.path.exists(ENGINE), reason="engine not built (needs the reference sources at build time)")


@needs_engine
@pytest.mark.parametrize("n,seed", [(2500, 11), (1500, 20261017)])
def test_own_board_planes_bit_identical(tmp_path, n, seed):
    out = str(tmp_path / "p.pos")
    subprocess.run([ENGINE, "-q", "--dump-planes", out, str(n), str(seed)], check=True, timeout=600)
    want, raw = fileio.read_positions(out), fileio.read_raw_positions(out + ".raw")
    assert want.n == n and raw.stones.shape == (n, 361)
    for i in range(n):
        pol, val = capi.planes_from_position(raw.stones[i], raw.to_move[i], raw.ko[i], raw.last[i], raw.prev[i], raw.komi[i])
        assert np.array_equal(pol, want.policy_planes[i]), f"policy planes differ at position {i} (move {want.movenum[i]})"
        assert np.array_equal(val, want.value_planes[i]), f"value planes differ at position {i} (move {want.movenum[i]})"
    # the set exercises what is hard: ladders (both kinds), captures, kos, full boards
    assert (want.policy_planes >> 25 & 1).sum() > 50 and (want.policy_planes >> 26 & 1).sum() > 300
    assert (want.policy_planes >> 27 & 1).sum() > 0 and want.movenum.max() > 300


def test_hand_made_positions():
    empty = np.zeros(361, np.uint8)
    pol, val = capi.planes_from_position(empty, 0)
    line3 = np.array([(i % 19 in (2, 16)) or (i // 19 in (2, 16)) for i in range(361)])
    assert ((pol >> 0) & 1).all() and ((pol >> 31) & 1).astype(bool).tolist() == line3.tolist()
    assert (((pol >> 13) & 0x3f) != 0).all()           # every point has a liberties-after-move plane set
    corner_after = (pol[0] >> 13) & 0x3f
    assert corner_after == 0b10                         # a corner stone has 2 liberties
    assert ((val >> 30) & 1).astype(bool).tolist() == line3.tolist()
    # a working ladder: white stone at (3,3) in atari after black's net of stones; black to move can capture in a ladder
    st = empty.copy()
    def at(x, y): return y * 19 + x
    st[at(3, 3)] = 2
    for x, y in ((2, 3), (3, 2), (4, 4)):
        st[at(x, y)] = 1
    pol, _ = capi.planes_from_position(st, 0, komi=7.5)
    assert (pol[at(3, 3)] >> 2) & 1 and (pol[at(3, 3)] >> 9) & 1      # opponent stone with 2 liberties
    assert (pol[at(3, 4)] >> 26) & 1 or (pol[at(4, 3)] >> 26) & 1      # one of the ataris starts a winning ladder
    assert (pol[at(3, 3)] >> 30) & 1                                   # white stones carry the komi plane
    pol0, _ = capi.planes_from_position(st, 0, komi=0.5)
    assert not (pol0[at(3, 3)] >> 30) & 1
    # history and ko planes
    pol, val = capi.planes_from_position(st, 1, ko_point=at(10, 10), last_move=at(4, 4), prev_move=at(3, 3))
    assert (pol[at(10, 10)] >> 27) & 1 and (pol[at(4, 4)] >> 28) & 1 and (pol[at(3, 3)] >> 29) & 1
    assert (val[at(10, 10)] >> 31) & 1
    pol, _ = capi.planes_from_position(st, 1, last_move=-1, prev_move=at(3, 3))   # no last move: no history at all
    assert not ((pol >> 28) & 3).any()


def test_bad_arguments_are_rejected():
    st = np.zeros(361, np.uint8)
    st[5] = 3
    with pytest.raises(capi.Lb2Error) as e:
        capi.planes_from_position(st, 0)
    assert e.value.code == -1
    with pytest.raises(capi.Lb2Error):
        capi.planes_from_position(np.zeros(361, np.uint8), 0, ko_point=361)


def _sym_index(s):
    """where board point idx lands under symmetry s of the square (bit 2: transpose, bit 0: flip y, bit 1: flip x)"""
    out = np.empty(361, np.int64)
    for idx in range(361):
        x, y = idx % 19, idx // 19
        if s & 4: x, y = y, x
        if s & 1: y = 18 - y
        if s & 2: x = 18 - x
        out[idx] = y * 19 + x
    return out


@needs_engine
def test_planes_commute_with_board_symmetries_and_colour_swap(tmp_path):
    """Size-independent properties of the path: liberties, liberties after a move and both ladder readers do not know
    directions or colours. (1) Mapping a position through any of the 8 symmetries of the board permutes its planes the same
    way (the third-line plane is symmetric itself). (2) Swapping the colours of all stones AND the side to move leaves every
    plane where it is, except the white-has-komi plane, which follows the white stones."""
    out = str(tmp_path / "p.pos")
    subprocess.run([ENGINE, "-q", "--dump-planes", out, "400", "4242"], check=True, timeout=600)
    raw = fileio.read_raw_positions(out + ".raw")
    komi_plane_p, komi_plane_v = 1 << 30, 1 << 29
    checked_ladders = 0
    for i in range(0, 400, 2):
        st, tm, ko, last, prev, komi = raw.stones[i], int(raw.to_move[i]), int(raw.ko[i]), int(raw.last[i]), int(raw.prev[i]), float(raw.komi[i])
        pol, val = capi.planes_from_position(st, tm, ko, last, prev, komi)
        checked_ladders += int(((pol >> 25) & 3).any())
        for s in (1, 2, 4, 7, 5):
            to = _sym_index(s)
            st2 = np.zeros_like(st); st2[to] = st
            mv = lambda v: int(to[v]) if v >= 0 else v
            pol2, val2 = capi.planes_from_position(st2, tm, mv(ko), mv(last), mv(prev), komi)
            want_p = np.zeros_like(pol); want_p[to] = pol
            want_v = np.zeros_like(val); want_v[to] = val
            assert np.array_equal(pol2, want_p) and np.array_equal(val2, want_v), (i, s)
        swapped = np.where(st == 1, 2, np.where(st == 2, 1, 0)).astype(st.dtype)
        pol3, val3 = capi.planes_from_position(swapped, 1 - tm, ko, last, prev, komi)
        assert np.array_equal(pol3 & ~np.uint32(komi_plane_p), pol & ~np.uint32(komi_plane_p)), i
        assert np.array_equal(val3 & ~np.uint32(komi_plane_v), val & ~np.uint32(komi_plane_v)), i
        if abs(komi) > 0.75:
            assert np.array_equal((pol3 & komi_plane_p) != 0, swapped == 2) and np.array_equal((val3 & komi_plane_v) != 0, swapped == 2)
    assert checked_ladders > 20   # the sample has positions where the ladder readers fire


def test_random_boards_satisfy_the_plane_invariants():
    """Arbitrary stone patterns through the C ABI — including ones no game reaches (strings without liberties, full boards):
    the builder must return (bounded ladder reads, no out-of-range string ids) and its planes must be well-formed: every point
    is exactly one of empty / own / opponent, a stone carries at most one liberty-count plane of its own side, and the
    after-a-move and ladder planes only ever sit on empty points."""
    rng = np.random.default_rng(2026)
    t0 = __import__("time").time()
    for trial in range(300):
        density = rng.uniform(0.05, 0.98)
        r = rng.random(361)
        st = np.where(r < density / 2, 1, np.where(r < density, 2, 0)).astype(np.uint8)
        tm = int(rng.integers(0, 2))
        empties = np.flatnonzero(st == 0)
        ko = int(rng.choice(empties)) if len(empties) and rng.random() < 0.3 else -1
        stones = np.flatnonzero(st != 0)
        last = int(rng.choice(stones)) if len(stones) and rng.random() < 0.8 else -1
        prev = int(rng.choice(stones)) if len(stones) and rng.random() < 0.8 else -1
        pol, val = capi.planes_from_position(st, tm, ko, last, prev, float(rng.choice([0.5, 7.5, -7.5])))
        own, opp = (1, 2) if tm == 0 else (2, 1)
        for planes, libs_own, libs_opp, n_lib, after_lo, after_hi, ladders in ((pol, 3, 8, 5, 13, 24, (25, 26)), (val, 3, 9, 6, 15, 26, (27, 28))):
            p = planes.astype(np.uint64)
            assert np.array_equal((p >> 0) & 1, st == 0) and np.array_equal((p >> 1) & 1, st == own) and np.array_equal((p >> 2) & 1, st == opp)
            own_bits = (p >> libs_own) & ((1 << n_lib) - 1)
            opp_bits = (p >> libs_opp) & ((1 << n_lib) - 1)
            popcount = lambda x: np.array([bin(int(v)).count("1") for v in x])
            assert (popcount(own_bits)[st == own] <= 1).all() and (own_bits[st != own] == 0).all()
            assert (popcount(opp_bits)[st == opp] <= 1).all() and (opp_bits[st != opp] == 0).all()
            after = (p >> after_lo) & ((1 << (after_hi - after_lo + 1)) - 1)
            assert (after[st != 0] == 0).all()
            for b in ladders:
                assert (((p >> b) & 1)[st != 0] == 0).all()
    assert __import__("time").time() - t0 < 60
