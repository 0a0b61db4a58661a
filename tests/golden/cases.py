"""Deterministic inputs of the per-layer golden cases (shared by make_golden.py and the tests).
Built on the counter-based generator in leela_b200.synth, so they do not depend on numpy's RNG."""
import numpy as np

from leela_b200 import synth
from leela_b200.netdefs import POLICY_CONVS, VALUE_CONVS

CONV_SHAPES = sorted({(c.k, c.c_in, c.c_out) for c in POLICY_CONVS + VALUE_CONVS})
IP_SHAPES = ((361, 256), (256, 1))


def conv_case(k, ci, co, seed=7):
    """x [ci,361] in [-1,1) (fp16-representable), w OIHW (fp16-representable), b [co]."""
    tag = 1000 + k * 100000 + ci * 300 + co
    x = synth._unit(seed, tag, ci * 361).astype(np.float16).astype(np.float32).reshape(ci, 361)
    w = synth.synth_weights(co * ci * k * k, seed, tag + 1, ci * k * k)
    w = w.astype(np.float16).astype(np.float32).reshape(co, ci, k, k)
    b = synth.synth_biases(co, seed, tag + 2)
    return x, w, b


def ip_case(ni, no, seed=7):
    tag = 5000 + ni * 300 + no
    x = synth._unit(seed, tag, ni).astype(np.float32)
    w = synth.synth_weights(no * ni, seed, tag + 1, ni).reshape(no, ni)
    b = synth.synth_biases(no, seed, tag + 2)
    return x, w, b
