"""Builds the drop-in engine (engine/_build/leela_b200_engine) and writes the synthetic weights
file it loads. The reference's sources are compiled where they lie (engine/Makefile), so this
only works where /root/reference exists (the build container); the binary then travels."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFERENCE_SRC = "/root/reference"
ENGINE = os.path.join(HERE, "_build", "leela_b200_engine")
WEIGHTS = os.path.join(HERE, "_build", "weights_synth.lb2w")


def synth_kat():
    """Known answers for the synthetic weights: what the REFERENCE's OpenBLAS path returns on the empty board
    (tests/golden/ref_golden.npz position 0: move 0, black to move, symmetry 0) at its three most probable points."""
    import numpy as np
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_golden.npz"))
    assert int(g["movenum"][0]) == 0 and int(g["rotation"][0]) == 0
    p = g["policy"][0]
    top = np.argsort(-p)[:3]
    # on the empty board Netresult entry i is board point i; FastBoard vertex = (x + 1) + (y + 1) * 21
    return {"policy": [(int(i), int(i % 19 + 1 + (i // 19 + 1) * 21), float(p[i])) for i in top], "value": float(g["value"][0])}


def write_synth_weights(path: str = WEIGHTS) -> str:
    sys.path.insert(0, ROOT)
    from leela_b200 import fileio, synth
    os.makedirs(os.path.dirname(path), exist_ok=True)
    fileio.write_weights(path, {0: synth.policy_weights(), 1: synth.value_weights()}, kat=synth_kat())
    return path


def build() -> bool:
    if os.path.isdir(REFERENCE_SRC):
        subprocess.check_call(["make", "-C", HERE, "-j", str(os.cpu_count() or 4)], stdout=subprocess.DEVNULL)
    write_synth_weights()
    return os.path.exists(ENGINE)


if __name__ == "__main__":
    print("engine built:", build())
