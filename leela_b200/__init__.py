"""leela_b200 — B200 (sm_100a) evaluator for Leela's policy and value networks.

The product is the C-ABI shared library (include/leela_b200.h, leela_b200/csrc); this Python
package is the thin binding used by tests and bench.py plus the layer geometry and the synthetic
weight generator (the reference's weight files are missing from the snapshot).
"""
from . import netdefs, synth  # noqa: F401
