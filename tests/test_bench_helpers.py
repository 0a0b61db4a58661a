"""bench.py's bookkeeping that needs no GPU: the executed-work accounting beside the algorithmic flops, the kernel source
hash that ties profiles/trunk_traffic.json to a build, and the flags a multi-GPU relaunch has to forward."""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from leela_b200 import netdefs  # noqa: E402


def test_algorithmic_flops_are_the_survey_figures():
    assert netdefs.POLICY_FLOPS == 1_200_761_088 and netdefs.VALUE_FLOPS == 303_725_696
    # the trunk kernel's share: everything but the two C -> 1 convs and the inner products
    assert bench.TRUNK_FLOPS == netdefs.POLICY_FLOPS + netdefs.VALUE_FLOPS - 2 * 361 * 9 * (128 + 64) - 2 * (361 * 256 + 256)


def test_executed_work_accounting():
    """padded row space (400 rows per 361-point position, 441 in front of the 5x5 layer) x the K-loop terms of the mode"""
    algo = bench.TRUNK_FLOPS * 256
    fp16 = bench.executed_trunk_flops((0, 0), 256)
    assert 1.10 < fp16 / algo < 1.13
    default = bench.executed_trunk_flops((0, 1), 256)
    # the value net's K loop runs twice in lite mode: + its padded work once more
    value_padded = sum(2 * (-(-256 * (441 if c.k == 5 else 400) // 512) * 512) * c.k * c.k * c.c_in * c.c_out for c in netdefs.VALUE_CONVS[:-1])
    assert default - fp16 == value_padded
    assert 1.33 < default / algo < 1.35
    full = bench.executed_trunk_flops((2, 2), 256)
    assert 3.2 < full / algo < 3.3   # three terms, two for the layers with binary inputs
    # whole 512-row items: one position still costs a whole item per layer
    assert bench.executed_trunk_flops((0, 0), 1) == sum(2 * 512 * c.k * c.k * c.c_in * c.c_out for convs in (netdefs.POLICY_CONVS, netdefs.VALUE_CONVS) for c in convs[:-1])


def test_kernel_source_hash_and_traffic_record():
    sha = bench.kernel_source_sha()
    assert re.fullmatch(r"[0-9a-f]{16}", sha)
    with open(os.path.join(ROOT, "profiles", "trunk_traffic.json")) as f:
        tj = json.load(f)
    # the record names the build it was captured from; bench.py reports `traffic` only while the two agree
    assert re.fullmatch(r"[0-9a-f]{16}", tj["kernel_source_sha16"]) and tj["mode"] == [0, 1]
    assert tj["dram_bytes_per_launch"] == tj["dram_bytes_read"] + tj["dram_bytes_write"]


def test_relaunch_forwards_every_measurement_flag():
    """the torchrun relaunch for --gpus > 1 must carry every flag that changes what is measured"""
    src = open(os.path.join(ROOT, "bench.py")).read()
    relaunch = src[src.index("convenience: relaunch under torchrun"):]
    for flag in ("--steps", "--warmup", "--batch", "--e2e-threads", "--flush-l2", "--no-cpu", "--precise", "--no-graphs",
                 "--policy-precision", "--value-precision", "--opt"):
        assert flag in relaunch, flag


def test_trunk_model_agrees_with_the_executed_work_accounting():
    """tools/trunk_model.py derives the issued MMA work from the layer shapes; bench.py's `roofline.executed` must be the same figure"""
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "trunk_model.py"), "--measured-us", "369"], capture_output=True, text=True, check=True).stdout
    m = re.search(r"executed \(issued\) ([\d.]+) G bf16-equivalent = x([\d.]+)", out)
    assert m, out
    assert abs(float(m.group(1)) * 1e9 - bench.executed_trunk_flops((0, 1), 256)) / bench.executed_trunk_flops((0, 1), 256) < 1e-3
    assert re.search(r"0\.8\d of the formulation bound", out), out
