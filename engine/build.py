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


def write_synth_weights(path: str = WEIGHTS) -> str:
    sys.path.insert(0, ROOT)
    from leela_b200 import fileio, synth
    os.makedirs(os.path.dirname(path), exist_ok=True)
    fileio.write_weights(path, {0: synth.policy_weights(), 1: synth.value_weights()})
    return path


def build() -> bool:
    if os.path.isdir(REFERENCE_SRC):
        subprocess.check_call(["make", "-C", HERE, "-j", str(os.cpu_count() or 4)], stdout=subprocess.DEVNULL)
    if not os.path.exists(WEIGHTS):
        write_synth_weights()
    return os.path.exists(ENGINE)


if __name__ == "__main__":
    print("engine built:", build())
