"""CPU-side checks of the boundary: the shared library loads and exports every symbol that
include/leela_b200.h declares; without a GPU lb2_init fails with an error code (no fallback)."""
import ctypes
import os
import re

import pytest

from leela_b200 import build, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    return build.build()


def test_header_symbols_are_exported(lib_path):
    hdr = open(os.path.join(ROOT, "include", "leela_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(lb2_[a-z0-9_]+)\s*\(", hdr)) - {"lb2_callback"})
    assert declared, "no declarations parsed"
    L = ctypes.CDLL(lib_path)
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(capi.EXPORTS) == declared


def test_no_cpu_fallback(lib_path):
    """Without a B200 the product path must fail loudly, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.Lb2Error):
        capi.Evaluator()


def test_product_does_not_import_oracle():
    """The product path must never import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "leela_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle|liboracle|ref_harness\b(?!\.cpp)|oracle/_ref", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), f"{f} uses oracle/"
