"""Readers/writers for the two flat binary formats exchanged with the reference harness
(oracle/ref/ref_harness.cpp): positions ("LB2POS01") and evaluation outputs ("LB2OUT01").

Positions carry, per board position, the 32 policy planes and the 32 value planes of
Network::gather_features_policy / _value (Network.cpp:883-1201) packed one uint32 per board
point (idx = y*19 + x, bit c = plane c), plus the symmetry (0..7) to evaluate it under.
"""
from __future__ import annotations

import dataclasses
import numpy as np

P = 361


@dataclasses.dataclass
class Positions:
    policy_planes: np.ndarray  # uint32 [n, 361]
    value_planes: np.ndarray   # uint32 [n, 361]
    rotation: np.ndarray       # uint8  [n]
    to_move: np.ndarray        # int32  [n]
    movenum: np.ndarray        # int32  [n]

    @property
    def n(self) -> int:
        return int(self.policy_planes.shape[0])

    def take(self, idx) -> "Positions":
        return Positions(*(getattr(self, f.name)[idx] for f in dataclasses.fields(self)))


@dataclasses.dataclass
class RefOutputs:
    softmax_temp: float
    policy: np.ndarray      # float32 [n, 361], un-rotated, all 361 points
    value: np.ndarray       # float32 [n]
    policy_avg: np.ndarray  # float32 [n_avg, 361]; AVERAGE_ALL via the API, -1 where not EMPTY
    value_avg: np.ndarray   # float32 [n_avg]


def read_positions(path) -> Positions:
    raw = np.fromfile(path, dtype=np.uint8)
    if raw[:8].tobytes() != b"LB2POS01":
        raise ValueError(f"{path}: not a positions file")
    n = int(raw[8:12].view(np.int32)[0])
    off = 16
    def take(count, dtype):
        nonlocal off
        nbytes = count * np.dtype(dtype).itemsize
        a = raw[off:off + nbytes].view(dtype).copy()
        off += nbytes
        return a
    pol = take(n * P, np.uint32).reshape(n, P)
    val = take(n * P, np.uint32).reshape(n, P)
    rot = take((n + 3) // 4 * 4, np.uint8)[:n]
    tm = take(n, np.int32)
    mv = take(n, np.int32)
    return Positions(pol, val, rot, tm, mv)


def write_positions(path, ps: Positions) -> None:
    n = ps.n
    with open(path, "wb") as f:
        f.write(b"LB2POS01")
        f.write(np.array([n, 0], dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(ps.policy_planes, dtype=np.uint32).tobytes())
        f.write(np.ascontiguousarray(ps.value_planes, dtype=np.uint32).tobytes())
        rot = np.zeros((n + 3) // 4 * 4, dtype=np.uint8)
        rot[:n] = ps.rotation
        f.write(rot.tobytes())
        f.write(np.ascontiguousarray(ps.to_move, dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(ps.movenum, dtype=np.int32).tobytes())


def read_outputs(path) -> RefOutputs:
    raw = np.fromfile(path, dtype=np.uint8)
    if raw[:8].tobytes() != b"LB2OUT01":
        raise ValueError(f"{path}: not an outputs file")
    n, n_avg = (int(v) for v in raw[8:16].view(np.int32))
    temp = float(raw[16:20].view(np.float32)[0])
    body = raw[20:].view(np.float32)
    o = 0
    pol = body[o:o + n * P].reshape(n, P).copy(); o += n * P
    val = body[o:o + n].copy(); o += n
    pavg = body[o:o + n_avg * P].reshape(n_avg, P).copy(); o += n_avg * P
    vavg = body[o:o + n_avg].copy()
    return RefOutputs(temp, pol, val, pavg, vavg)


def write_weights(path, nets, kat=None) -> None:
    """Weights file of the drop-in engine (engine/network_b200.cpp:load_weights), "LB2WGT01".

    `kat` (optional): known answers the engine checks at start-up, as the reference's OpenCL self-test does for its own
    weights (GTP.cpp:105-125): {"policy": [(netresult index, vertex, probability), ...], "value": winrate} for
    Network::get_scored_moves(DIRECT, 0) / get_value(DIRECT, 0) on the empty 19x19 board — written as a trailer
    "LB2KAT01", count, {int32 index, int32 vertex, float32 p} x count, float32 value.

    `nets`: {kind: NetWeights} with kind 0 = policy, 1 = value. Per net: the conv arrays in OIHW
    order with their biases, then the inner products [n_out][n_in] — the layout of the reference's
    extern weight arrays (Network.cpp:54-137), in the order Network::initialize pushes them."""
    with open(path, "wb") as f:
        f.write(b"LB2WGT01")
        f.write(np.array([len(nets)], dtype=np.int32).tobytes())
        for kind in sorted(nets):
            w = nets[kind]
            f.write(np.array([kind, len(w.convs), len(w.ips)], dtype=np.int32).tobytes())
            for c, cw, cb in zip(w.convs, w.conv_w, w.conv_b):
                f.write(np.array([c.k, c.c_in, c.c_out], dtype=np.int32).tobytes())
                f.write(np.ascontiguousarray(cw, dtype=np.float32).tobytes())
                f.write(np.ascontiguousarray(cb, dtype=np.float32).tobytes())
            for p, pw, pb in zip(w.ips, w.ip_w, w.ip_b):
                f.write(np.array([p.n_in, p.n_out], dtype=np.int32).tobytes())
                f.write(np.ascontiguousarray(pw, dtype=np.float32).tobytes())
                f.write(np.ascontiguousarray(pb, dtype=np.float32).tobytes())
        if kat:
            f.write(b"LB2KAT01")
            f.write(np.array([len(kat["policy"])], dtype=np.int32).tobytes())
            for index, vertex, prob in kat["policy"]:
                f.write(np.array([index, vertex], dtype=np.int32).tobytes())
                f.write(np.array([prob], dtype=np.float32).tobytes())
            f.write(np.array([kat["value"]], dtype=np.float32).tobytes())


@dataclasses.dataclass
class RawPositions:
    """What lb2_planes_from_position takes (engine --dump-planes writes OUT.raw, "LB2RAW01")."""
    stones: np.ndarray     # uint8 [n, 361]: 0 empty, 1 black, 2 white
    to_move: np.ndarray    # int32 [n]: 0 black, 1 white
    ko: np.ndarray         # int32 [n]: idx or -1
    last: np.ndarray       # int32 [n]: idx or -1 (none / pass)
    prev: np.ndarray       # int32 [n]
    komi: np.ndarray       # float32 [n]


def read_raw_positions(path) -> RawPositions:
    raw = np.fromfile(path, dtype=np.uint8)
    if raw[:8].tobytes() != b"LB2RAW01":
        raise ValueError(f"{path}: not a raw positions file")
    n = int(raw[8:12].view(np.int32)[0])
    off = 16
    stones = raw[off:off + n * P].reshape(n, P).copy(); off += n * P
    ints = raw[off:off + 16 * n].view(np.int32).reshape(4, n).copy(); off += 16 * n
    komi = raw[off:off + 4 * n].view(np.float32).copy()
    return RawPositions(stones, ints[0], ints[1], ints[2], ints[3], komi)
