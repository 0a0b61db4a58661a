#!/usr/bin/env python
"""bench.py — positions/s of batched policy+value evaluation (BASELINE.json's metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch 256] [--impl ours|reference]

A step = one pass of the hot path over one batch of 256 synthetic positions per GPU (policy net
NN128 + value net NNValue shapes, seeded synthetic weights — the reference's weight files are
missing from the snapshot — on positions from Leela's own Playout self-play, tests/golden/
bench_positions.npz). N > 1 is launched by torchrun, one rank per GPU; weights replicated,
positions sharded, no collective on the data path (weak scaling).

  value     kernels only: inputs resident in HBM, CUDA events around every step on the launching
            stream, max over ranks. Every step reads a different input batch out of a pool larger
            than L2 (192 MB of packed planes, 260 batches); weights and the activation workspace are
            re-used from step to step exactly as in steady-state serving. --flush-l2 instead
            evicts L2 (256 MB write, outside the timed region) before every step.
  e2e       the same metric through the C ABI with HOST buffers (lb2_eval_both): pinned host
            planes -> H2D -> kernels -> D2H of probabilities and winrates inside the timed region,
            two host threads calling concurrently (the search's threads do), so the copies of one
            call overlap the kernels of the other; --e2e-threads 1 gives the single-caller figure.
  roofline  trunk_kernel (tcgen05 conv stack): algorithmic FLOPs per launch / its CUDA-event time,
            against the measured cuBLAS bf16 peak SUSTAINED under the power cap (the kernel is timed
            inside a long back-to-back run with sw_power_cap active); the burst figure is reported too.
  cpu_baseline  the reference's own OpenBLAS path (oracle/_ref, built from /root/reference) on all
            host cores, bounded sample, rank 0 at N=1 only.

--impl reference times that CPU path alone with the same JSON shape.
--engine runs the engine-level benchmark instead (GTP genmove at a fixed think time and netbench:
the drop-in engine beside the reference's own CPU engine); it is not part of the driver's contract.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from leela_b200 import fileio, netdefs, shard  # noqa: E402

METRIC = "nn_evals_per_sec_policy_plus_value_batch256"
UNIT = "positions/s"
TEMP = 0.75
TRUNK_FLOPS = (netdefs.POLICY_FLOPS - 2 * 361 * 9 * 128) + \
              (netdefs.VALUE_FLOPS - 2 * 361 * 9 * 64 - 2 * (361 * 256 + 256))  # per position, heads excluded


def load_positions():
    g = np.load(os.path.join(ROOT, "tests", "golden", "bench_positions.npz"))
    return g["policy_planes"], g["value_planes"], g["rotation"]


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except OSError:
        return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(prefix="lb2clk_", suffix=".csv")
        self.f = open(self.path, "w")
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(gpu_index), "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.p.kill()
        self.f.close()
        sm, mx, pw, reasons = [], [], [], set()
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1])); mx.append(float(c[2]))
                except ValueError:
                    continue
                try:
                    pw.append(float(c[3]))
                except ValueError:
                    pass
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        top = sorted(sm)[len(sm) // 2:]  # samples under load = upper half
        return {"sm_mhz": statistics.median(top), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(pw) if pw else None}


def cpu_reference(sample, warmup, steps, threads=None):
    """Times the reference's own CPU path (oracle/_ref) — or, if it is unavailable on this box,
    the plain-C port — on `threads` host cores. Returns (pos_per_s, dict)."""
    from oracle import reference
    threads = threads or (os.cpu_count() or 1)
    pp, vp, rot = load_positions()
    if reference.available():
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "bench.pos")
            n = pp.shape[0]
            fileio.write_positions(path, fileio.Positions(pp, vp, rot, np.zeros(n, np.int32), np.zeros(n, np.int32)))
            r = reference.steps(path, threads, sample, warmup, steps, "both")
        return r["pos_per_s"], {"kind": "reference", "cores": threads, "blas_core": r["blas_core"],
                                "sample": f"{steps} steps x {sample} positions (policy+value, batch 1 per thread as "
                                          f"Network::benchmark), OpenBLAS 1 thread/worker"}
    from leela_b200 import synth
    from oracle import oracle
    pn, vn = oracle.OracleNet(synth.policy_weights()), oracle.OracleNet(synth.value_weights())
    k = min(sample, 4 * threads)
    t0 = time.perf_counter()
    for s in range(steps):
        sl = slice((s * k) % 512, (s * k) % 512 + k)
        oracle.policy_forward(pn, pp[sl], rot[sl], TEMP)
        oracle.value_forward(vn, vp[sl], rot[sl])
    dt = time.perf_counter() - t0
    return k * steps / dt, {"kind": "port", "cores": min(threads, 32),
                            "sample": f"{steps} steps x {k} positions through oracle/leela_oracle.c"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    sample = args.batch
    # keep the whole run within a few minutes: ~30 positions/s/core expected
    est = (args.steps + args.warmup) * sample / (25.0 * threads)
    if est > 150:
        sample = max(threads, int(sample * 150 / est))
    v, info = cpu_reference(sample, args.warmup, args.steps, threads)
    info["value"] = v
    info["unit"] = UNIT
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sample / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"policy NN128 + value NNValue, {sample} positions per step on host cores "
                                   f"(reference evaluates batch 1 per thread)", "batch_per_step": sample},
            "cpu_baseline": info,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def run_engine(args):
    """BASELINE.json configs[4] at engine level: GTP genmove at a fixed think time (and `netbench`),
    the drop-in engine (reference search + B200 evaluator) beside the reference's own CPU engine
    (oracle/_ref/ref_engine, built from /root/reference). One JSON object per engine and leg."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import engine_bench as eb
    from oracle import reference
    ref_engine = os.path.join(ROOT, "oracle", "_ref", "ref_engine")
    env = reference._env()
    cores = os.cpu_count() or 1
    for netbench in (False, True):
        a = eb.parser().parse_args(["--threads", str(64 if netbench else min(cores, 64)), "--gpus", str(args.gpus)] +
                                   (["--netbench"] if netbench else []))
        print(json.dumps(eb.run_ours(a)), flush=True)
        threads = min(cores, 64)
        cmd = [ref_engine, "-g", "-t", str(threads), "--noponder", "--nobook", "--lagbuffer", "0"]
        print(json.dumps(eb.summarize("reference_cpu_engine", *eb.run(cmd, eb.script_for(a), env=env),
                                      {"threads": threads, "think_s": a.seconds, "blas_core": env.get("OPENBLAS_CORETYPE")})), flush=True)
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist
    from leela_b200 import capi, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200; there is no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    B = args.batch

    ev = capi.Evaluator(policy=synth.policy_weights(), value=synth.value_weights(), devices=[local])
    ev.set_option("max_batch", max(B, 256))
    if args.overlap_io:
        ev.set_option("overlap_io", 1)
    if args.precise:
        ev.set_option("precise", 1)
    pp, vp, rot = load_positions()
    n_pos = pp.shape[0]
    # input pool: every batch is a different window (stride 3, wrapping) over the distinct positions,
    # each in its own device buffer; together they exceed the 126 MB L2
    set_bytes = B * 361 * 4 * 2 + B
    n_sets = 4 if args.flush_l2 else max(4, -(-192 * 1024 * 1024 // set_bytes))
    first = [(3 * s + 17 * rank) % n_pos for s in range(n_sets)]
    window = [np.arange(f, f + B) % n_pos for f in first]
    d_pp = [torch.from_numpy(pp[w].astype(np.int32)).to(dev) for w in window]
    d_vp = [torch.from_numpy(vp[w].astype(np.int32)).to(dev) for w in window]
    d_rot = [torch.from_numpy(rot[w].copy()).to(dev) for w in window]
    d_probs = torch.empty((B, 361), dtype=torch.float32, device=dev)
    d_win = torch.empty((B,), dtype=torch.float32, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev) if args.flush_l2 else None  # > 126 MB L2
    # a dedicated (non-default) stream: kernels, L2 flush and the timing events all go on it
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)

    def step(i):
        s = i % n_sets
        ev.eval_both_device(d_pp[s].data_ptr(), d_vp[s].data_ptr(), d_rot[s].data_ptr(), B, TEMP,
                            d_probs.data_ptr(), d_win.data_ptr(), stream=stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------------------------------------------------------- kernels, inputs in HBM
    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ev.set_option("profile_trunk", 1)
    ev.get_option("trunk_ns")
    launches0 = ev.launch_count
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        if flush is not None:
            flush.fill_(i & 0xFF)        # evict L2 between steps; outside the timed region
        starts[i].record(stream)
        step(i)
        stops[i].record(stream)
    barrier()
    launches = ev.launch_count - launches0
    step_ms = [a.elapsed_time(b) for a, b in zip(starts, stops)]
    total_ms = sum(step_ms)
    trunk_ns = ev.get_option("trunk_ns")
    ev.set_option("profile_trunk", 0)

    # ---------------------------------------------------------------- end to end through the C ABI
    import threading
    n_host = max(1, args.e2e_threads)
    host_sets = min(n_sets, 16)
    h_pp = [torch.from_numpy(pp[w].astype(np.int32)).pin_memory() for w in window[:host_sets]]
    h_vp = [torch.from_numpy(vp[w].astype(np.int32)).pin_memory() for w in window[:host_sets]]
    h_rot = [torch.from_numpy(rot[w].copy()).pin_memory() for w in window[:host_sets]]
    h_probs = [torch.empty((B, 361), dtype=torch.float32).pin_memory() for _ in range(n_host)]
    h_win = [torch.empty((B,), dtype=torch.float32).pin_memory() for _ in range(n_host)]

    def step_e2e(i, t):
        s = i % host_sets
        ev.eval_both_raw(h_pp[s].data_ptr(), h_vp[s].data_ptr(), h_rot[s].data_ptr(), B, TEMP,
                         h_probs[t].data_ptr(), h_win[t].data_ptr())   # blocking: returns after the D2H of this step's results

    def caller(t, lo, hi):
        for i in range(lo, hi):
            step_e2e(i, t)

    for i in range(3):
        step_e2e(i, 0)
    barrier()
    bounds = [args.steps * t // n_host for t in range(n_host + 1)]
    threads = [threading.Thread(target=caller, args=(t, bounds[t], bounds[t + 1])) for t in range(n_host)]
    t0 = time.perf_counter()
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    checksum = float(h_probs[0].sum()) + float(h_win[0].sum())
    barrier()
    clocks = sampler.stop() if sampler else None

    total_ms, e2e_ms = shard.max_over_ranks([total_ms, e2e_s * 1e3], dist if world > 1 else None, dev)

    if rank == 0:
        n_gpus = world
        value = shard.aggregate_throughput(B, args.steps, n_gpus, total_ms)
        e2e_value = shard.aggregate_throughput(B, args.steps, n_gpus, e2e_ms)
        peaks, peak_src = measured_peaks()
        trunk_s = trunk_ns * 1e-9 / args.steps
        achieved = TRUNK_FLOPS * B / trunk_s / 1e12 if trunk_s > 0 else 0.0
        peak_burst = float(peaks["bf16_tflops"])
        peak = float(peaks.get("bf16_tflops_sustained", peak_burst))
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "trunk_traffic.json")) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except OSError:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16 hi+lo split operands (3 MMA terms), f32 accumulation" if args.precise else "f16", "data": "synthetic",
            "config": {"workload": "policy NN128 + value NNValue, batch 256 positions per GPU per step "
                                   "(BASELINE.json configs[1] shape, policy+value as the metric names)",
                       "batch_per_gpu": B, "global_batch": B * n_gpus, "parallelism": f"replicas x{n_gpus}, positions sharded",
                       "weights": "synthetic U(+-sqrt(6/fan_in)), seed 20260001, policy gain 2 (in-repo weights missing)",
                       "positions": f"Leela Playout self-play, {n_pos} distinct, {n_sets} different batches cycled",
                       "l2": "flushed between steps (256 MB write)" if args.flush_l2 else
                             f"inputs larger than L2: {n_sets} input batches = {n_sets * set_bytes / 2**20:.0f} MB, one per step; weights + workspace stay warm",
                       "trunk_mode": ev.get_option("trunk_mode"), "cta_pair": ev.get_option("cta_pair"), "precise": ev.get_option("precise"), "flops_per_position": netdefs.POLICY_FLOPS + netdefs.VALUE_FLOPS,
                       "pct_of_bf16_sustained_peak_whole_step": 100.0 * (netdefs.POLICY_FLOPS + netdefs.VALUE_FLOPS) * value / n_gpus / (peak * 1e12),
                       "pct_of_bf16_burst_peak_whole_step": 100.0 * (netdefs.POLICY_FLOPS + netdefs.VALUE_FLOPS) * value / n_gpus / (peak_burst * 1e12)},
            "roofline": {"bound": "tensor", "kernel": "trunk_kernel", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "peak_source": f"{peak_src} sustained cuBLAS bf16 under the power cap (MEASURED_PEAKS.json bf16_tflops_sustained): "
                                        "the kernel is timed inside a long back-to-back run with sw_power_cap active",
                         "peak_burst": peak_burst, "frac_of_burst": achieved / peak_burst,
                         "flops_per_launch": TRUNK_FLOPS * B, "launch_ms": trunk_s * 1e3},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * (2 * 1444 + 1),
                    "d2h_bytes_per_step": B * (1444 + 4), "timing": f"host clock around {n_host} host thread(s) each making blocking C-ABI calls", "host_threads": n_host, "checksum": checksum},
            "gpu_launches": launches, "clocks": clocks,
        }
        if n_gpus == 1 and not args.no_cpu:
            v, info = cpu_reference(min(B, 256), 1, 8)
            info["value"] = v
            info["unit"] = UNIT
            line["cpu_baseline"] = info
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    ev.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--engine", action="store_true", help="engine-level benchmark instead: GTP genmove / netbench, ours vs the reference's CPU engine")
    ap.add_argument("--flush-l2", action="store_true", help="evict L2 before every step instead of cycling an input pool larger than L2")
    ap.add_argument("--overlap-io", action="store_true", help="A/B: expand/heads kernels of host-buffer calls on the I/O slot's stream")
    ap.add_argument("--precise", action="store_true", help="split-operand mode (lb2_set_option precise=1): results within 1e-4 of the fp32 reference, "
                    "3x the tensor work; the roofline still counts the ALGORITHMIC flops")
    ap.add_argument("--e2e-threads", type=int, default=2, help="host threads calling the C ABI concurrently in the e2e leg")
    args = ap.parse_args()
    if args.engine:
        return run_engine(args)
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and "RANK" not in os.environ:
        # convenience: relaunch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__), "--gpus", str(args.gpus),
               "--steps", str(args.steps), "--warmup", str(args.warmup), "--batch", str(args.batch),
               "--e2e-threads", str(args.e2e_threads)] + (["--flush-l2"] if args.flush_l2 else []) + (["--no-cpu"] if args.no_cpu else [])
        return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
