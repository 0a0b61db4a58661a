"""leela_b200.synth (numpy) must be bit-identical to oracle/synth_weights.h (the generator the
reference-build harness uses for the missing NN128.cpp / NNValue.cpp arrays)."""
import os
import subprocess
import tempfile

import numpy as np

from leela_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

C_SRC = r'''
#include <stdio.h>
#include "synth_weights.h"
int main(void) {
    static float w[4096], b[64];
    lb2_synth_fill_weights(w, 4096, LB2_SYNTH_DEFAULT_SEED, 24, 1152, 2.0f);
    lb2_synth_fill_biases(b, 64, LB2_SYNTH_DEFAULT_SEED, 33);
    fwrite(w, 4, 4096, stdout); fwrite(b, 4, 64, stdout);
    return 0;
}
'''


def test_numpy_mirror_bit_identical():
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write(C_SRC)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-O2", "-I", os.path.join(ROOT, "oracle"), src, "-o", exe, "-lm"])
        raw = np.frombuffer(subprocess.check_output([exe]), dtype=np.float32)
    w = synth.synth_weights(4096, synth.DEFAULT_SEED, 24, 1152, 2.0)
    b = synth.synth_biases(64, synth.DEFAULT_SEED, 33)
    assert np.array_equal(raw[:4096].view(np.uint32), w.view(np.uint32))
    assert np.array_equal(raw[4096:].view(np.uint32), b.view(np.uint32))


def test_shapes():
    pw, vw = synth.policy_weights(), synth.value_weights()
    assert sum(w.size + b.size for w, b in zip(pw.conv_w, pw.conv_b)) == 1_664_609
    n = sum(w.size + b.size for w, b in zip(vw.conv_w, vw.conv_b)) + sum(w.size + b.size for w, b in zip(vw.ip_w, vw.ip_b))
    assert n == 514_050
