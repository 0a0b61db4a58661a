"""Regenerates tests/golden/*.npz by running the reference itself (oracle/_ref/ref_harness,
built from /root/reference by oracle/ref/Makefile). Run in the build container only:

    python tests/golden/make_golden.py

ref_golden.npz   96 self-play positions (position 0 = empty board, the reference's self-test
                 position): policy/value planes, rotation i%8, the reference's policy over all
                 361 points (get_scored_moves_internal on an empty-board filter) and value
                 (get_value_internal), plus AVERAGE_ALL through the public API
                 (get_scored_moves / get_value incl. EMPTY filter + ladder prune) for the
                 first 8; weights = leela_b200.synth defaults.
edge_golden.npz  hand-made edge cases (all-zero planes, all-one planes, single stones on
                 edges/corners, random bits) x all 8 rotations, evaluated by the reference.
layer_golden.npz one call of each convolve<> / innerproduct<> instantiation the nets use, on
                 the deterministic inputs of tests/golden/cases.py (large outputs strided by 5).
bench_positions.npz  1024 self-play positions (planes only) for bench.py.
bench_golden.npz the correctness set: the reference's policy (361 floats) and
                 value for all 1024 bench positions, rotation i%8 (`make_golden.py correctness`
                 regenerates only this file from the committed bench_positions.npz).
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from leela_b200 import fileio, synth  # noqa: E402
from oracle import reference  # noqa: E402
from tests.golden import cases  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def correctness_set():
    assert reference.build(), "reference harness unavailable"
    bp = np.load(os.path.join(HERE, "bench_positions.npz"))
    n = bp["policy_planes"].shape[0]
    ps = fileio.Positions(bp["policy_planes"], bp["value_planes"], bp["rotation"], np.zeros(n, np.int32), np.zeros(n, np.int32))
    out = reference.evaluate(ps)
    np.savez_compressed(os.path.join(HERE, "bench_golden.npz"), policy=out.policy, value=out.value,
                        softmax_temp=np.float32(out.softmax_temp))
    print("bench_golden.npz", os.path.getsize(os.path.join(HERE, "bench_golden.npz")))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "correctness":
        return correctness_set()
    assert reference.build(), "reference harness unavailable"
    with tempfile.TemporaryDirectory() as d:
        pre = os.path.join(d, "g")
        reference.dump(pre, 96, 1234, n_avg=8)
        ps = fileio.read_positions(pre + ".pos")
        out = fileio.read_outputs(pre + ".out")
        np.savez_compressed(os.path.join(HERE, "ref_golden.npz"),
                            policy_planes=ps.policy_planes, value_planes=ps.value_planes,
                            rotation=ps.rotation, to_move=ps.to_move, movenum=ps.movenum,
                            softmax_temp=np.float32(out.softmax_temp), policy=out.policy, value=out.value,
                            policy_avg=out.policy_avg, value_avg=out.value_avg,
                            weight_seed=np.int64(synth.DEFAULT_SEED), policy_gain=np.float32(synth.DEFAULT_POLICY_GAIN))

        # edge cases
        base = [np.zeros(361, np.uint32), np.full(361, 0xFFFFFFFF, np.uint32)]
        for idx in (0, 18, 342, 360, 9, 180):
            e = np.full(361, 1, np.uint32)           # plane 0 (empty) everywhere
            e[idx] = 0b010 | (1 << 28)               # one stone + "last move" plane there
            base.append(e)
        base += list(synth.random_planes(4, 99, 0.2))
        base = np.stack(base)
        reps = np.repeat(base, 8, axis=0)
        rot = np.tile(np.arange(8, dtype=np.uint8), base.shape[0])
        eps = fileio.Positions(reps, reps.copy(), rot, np.zeros(len(rot), np.int32), np.zeros(len(rot), np.int32))
        eo = reference.evaluate(eps)
        np.savez_compressed(os.path.join(HERE, "edge_golden.npz"), planes=reps, rotation=rot,
                            policy=eo.policy, value=eo.value, softmax_temp=np.float32(eo.softmax_temp))

        # per-layer (inputs come from tests/golden/cases.py; only outputs are stored)
        lay = {}
        for (k, ci, co) in cases.CONV_SHAPES:
            x, w, b = cases.conv_case(k, ci, co)
            y = reference.layer("conv", k, ci, co, x, w, b)
            lay[f"conv{k}_{ci}_{co}_y"] = y if y.size <= 8192 else y.reshape(-1)[::5].copy()
        for (ni, no) in cases.IP_SHAPES:
            x, w, b = cases.ip_case(ni, no)
            lay[f"ip_{ni}_{no}_y"] = reference.layer("ip", 0, ni, no, x, w, b)
        np.savez_compressed(os.path.join(HERE, "layer_golden.npz"), **lay)

        reference.planes(pre + "b", 1024, 777)
        bp = fileio.read_positions(pre + "b.pos")
        np.savez_compressed(os.path.join(HERE, "bench_positions.npz"), policy_planes=bp.policy_planes,
                            value_planes=bp.value_planes, rotation=bp.rotation)
    correctness_set()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
