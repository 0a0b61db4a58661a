"""Driver for oracle/_ref/ref_harness — the reference's own BLAS evaluation path compiled
from /root/reference (oracle/ref/Makefile). TEST INFRASTRUCTURE ONLY.

The binary is built in the build container (where /root/reference exists) and travels to the
GPU box as a prebuilt file; nothing here reads /root/reference at run time.
"""
from __future__ import annotations

import json
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
HARNESS = os.path.join(HERE, "_ref", "ref_harness")
REFERENCE_SRC = "/root/reference"


def available() -> bool:
    if not os.path.exists(HARNESS):
        return False
    try:
        r = subprocess.run([HARNESS], capture_output=True, timeout=20)
        return r.returncode == 2  # prints usage
    except Exception:
        return False


def build() -> bool:
    """(Re)build when the reference sources are present (build container only)."""
    if not os.path.isdir(REFERENCE_SRC):
        return available()
    subprocess.check_call(["make", "-C", os.path.join(HERE, "ref"), "-j", str(os.cpu_count() or 4)],
                          stdout=subprocess.DEVNULL)
    return available()


def _env(seed=None, gain=None):
    env = dict(os.environ)
    env.setdefault("OPENBLAS_NUM_THREADS", "1")
    # OpenBLAS 0.3.15 mis-detects recent Xeons and falls back to SSE kernels; pick the best
    # kernel the host supports (the reference's own match config sets OPENBLAS_CORETYPE too,
    # tools/gomill.ctl:16-19).
    if "OPENBLAS_CORETYPE" not in env:
        flags = ""
        try:
            with open("/proc/cpuinfo") as f:
                for line in f:
                    if line.startswith("flags"):
                        flags = line
                        break
        except OSError:
            pass
        if " avx512f" in flags and "GenuineIntel" in open("/proc/cpuinfo").read(4096):
            env["OPENBLAS_CORETYPE"] = "SkylakeX"
        elif " avx2" in flags:
            env["OPENBLAS_CORETYPE"] = "Haswell"
    if seed is not None:
        env["LB2_WEIGHT_SEED"] = str(seed)
    if gain is not None:
        env["LB2_POLICY_GAIN"] = repr(float(gain))
    return env


def run(args, seed=None, gain=None, timeout=3600) -> str:
    r = subprocess.run([HARNESS] + [str(a) for a in args], capture_output=True, text=True,
                       env=_env(seed, gain), timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"ref_harness {args} failed ({r.returncode}): {r.stderr[-2000:]}")
    return r.stdout


def dump(prefix, n, seed, n_avg=0, **kw):
    run(["dump", prefix, n, seed, n_avg], **kw)


def planes(prefix, n, seed, **kw):
    run(["planes", prefix, n, seed], **kw)


def evaluate(positions, **kw):
    """Reference outputs (policy [n,361] over all points, value [n]) for given planes."""
    from leela_b200 import fileio
    with tempfile.TemporaryDirectory() as d:
        pin, pout = os.path.join(d, "in.pos"), os.path.join(d, "out.bin")
        fileio.write_positions(pin, positions)
        run(["eval", pin, pout], **kw)
        return fileio.read_outputs(pout)


def layer(kind, k, c_in, c_out, x, w, b) -> np.ndarray:
    """One layer through the reference's convolve<> / innerproduct<> templates."""
    with tempfile.TemporaryDirectory() as d:
        paths = [os.path.join(d, f) for f in ("x", "w", "b", "o")]
        for p, a in zip(paths, (x, w, b)):
            np.ascontiguousarray(a, dtype=np.float32).tofile(p)
        run(["layer", kind, k, c_in, c_out] + paths)
        out = np.fromfile(paths[3], dtype=np.float32)
    return out.reshape(c_out, 361) if kind == "conv" else out


def bench(pos_path, threads, seconds, which="both", **kw) -> dict:
    return json.loads(run(["bench", pos_path, threads, seconds, which], timeout=seconds * 4 + 120, **kw))


def steps(pos_path, threads, sample, warmup, steps_, which="both", **kw) -> dict:
    return json.loads(run(["steps", pos_path, threads, sample, warmup, steps_, which], timeout=3600, **kw))
