"""Seeded synthetic weights with the reference's exact layer shapes.

The reference's weight files (NN128.cpp, NNValue.cpp) are missing from the snapshot
(.MISSING_LARGE_BLOBS), so tests and bench run on synthetic weights. This is the numpy mirror
of oracle/synth_weights.h (the generator the reference-build harness uses); the two produce
bit-identical fp32 arrays (tests/test_synth.py).
"""
from __future__ import annotations

import dataclasses
import numpy as np

from .netdefs import POLICY_CONVS, POLICY192_CONVS, VALUE_CONVS, VALUE_IPS

DEFAULT_SEED = 20260001
DEFAULT_POLICY_GAIN = 2.0
_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix(seed: int, arr_id: int, n: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        base = (np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
                + np.uint64(arr_id + 1) * np.uint64(0xD1B54A32D192ED03))
        z = base + np.arange(n, dtype=np.uint64)
        z ^= z >> np.uint64(30); z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27); z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    return z


def _unit(seed: int, arr_id: int, n: int) -> np.ndarray:
    u = (_mix(seed, arr_id, n) >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return np.float32(2.0) * u - np.float32(1.0)


def synth_weights(n: int, seed: int, arr_id: int, fan_in: int, gain: float = 1.0) -> np.ndarray:
    scale = np.float32(np.sqrt(6.0 / float(fan_in))) * np.float32(gain)
    return (_unit(seed, arr_id, n) * scale).astype(np.float32)


def synth_biases(n: int, seed: int, arr_id: int) -> np.ndarray:
    return (_unit(seed, arr_id, n) * np.float32(0.1)).astype(np.float32)


@dataclasses.dataclass
class NetWeights:
    """conv_w[l]: OIHW float32 [c_out, c_in, k, k]; conv_b[l]: [c_out]; ip_w[j]: [n_out, n_in]."""
    convs: tuple
    conv_w: list
    conv_b: list
    ips: tuple = ()
    ip_w: list = dataclasses.field(default_factory=list)
    ip_b: list = dataclasses.field(default_factory=list)


def policy_weights(seed: int = DEFAULT_SEED, gain: float = DEFAULT_POLICY_GAIN) -> NetWeights:
    w, b = [], []
    for i, c in enumerate(POLICY_CONVS):
        g = gain if i == len(POLICY_CONVS) - 1 else 1.0
        w.append(synth_weights(c.n_weights, seed, 2 * i, c.fan_in, g).reshape(c.c_out, c.c_in, c.k, c.k))
        b.append(synth_biases(c.c_out, seed, 2 * i + 1))
    return NetWeights(POLICY_CONVS, w, b)


def policy192_weights(seed: int = DEFAULT_SEED, gain: float = DEFAULT_POLICY_GAIN) -> NetWeights:
    """The 192-wide policy stack of the reference's OpenCL build (Network.cpp:55-80); array ids 64.."""
    w, b = [], []
    for i, c in enumerate(POLICY192_CONVS):
        g = gain if i == len(POLICY192_CONVS) - 1 else 1.0
        w.append(synth_weights(c.n_weights, seed, 64 + 2 * i, c.fan_in, g).reshape(c.c_out, c.c_in, c.k, c.k))
        b.append(synth_biases(c.c_out, seed, 65 + 2 * i))
    return NetWeights(POLICY192_CONVS, w, b)


def value_weights(seed: int = DEFAULT_SEED) -> NetWeights:
    w, b = [], []
    for j, c in enumerate(VALUE_CONVS):
        w.append(synth_weights(c.n_weights, seed, 32 + 2 * j, c.fan_in).reshape(c.c_out, c.c_in, c.k, c.k))
        b.append(synth_biases(c.c_out, seed, 33 + 2 * j))
    ipw = [synth_weights(361 * 256, seed, 56, 361).reshape(256, 361),
           synth_weights(256, seed, 58, 256).reshape(1, 256)]
    ipb = [synth_biases(256, seed, 57), synth_biases(1, seed, 59)]
    return NetWeights(VALUE_CONVS, w, b, VALUE_IPS, ipw, ipb)


def random_planes(n: int, seed: int, density: float = 0.25) -> np.ndarray:
    """Random bit-planes (uint32 [n, 361]); not Go positions, for shape/edge-case tests."""
    rng = np.random.default_rng(seed)
    bits = rng.random((n, 361, 32)) < density
    return (bits.astype(np.uint64) << np.arange(32, dtype=np.uint64)).sum(-1).astype(np.uint32)
