"""Engine-level benchmark (BASELINE.json configs[4]): GTP genmove at a fixed think time, playouts/s.

Runs engine/_build/leela_b200_engine — the reference's search (its own sources) with the B200
evaluator behind Network; search threads' single-position requests are coalesced into device batches
by the library's queue — through a GTP script and prints one JSON object.

Usage (GPU box): python tools/engine_bench.py [--seconds 5] [--moves 4] [--threads 16] [--gpus 1]
                 [--max-outstanding 2] [--netbench] [--extra "..."]
The same script against the reference's own CPU engine (the baseline) is `python bench.py --engine`:
only bench.py and tests/ may execute anything under oracle/.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "engine", "_build", "leela_b200_engine")
WEIGHTS = os.path.join(ROOT, "engine", "_build", "weights_synth.lb2w")
STATS = re.compile(r"(\d+) visits, (\d+) nodes, (\d+) playouts, (\d+) p/s")
BATCH = re.compile(r"B200 evaluator: (\d+) positions in (\d+) device batches \(mean batch ([\d.]+)\), (\d+) requests")


def gtp_script(seconds: int, moves: int) -> str:
    lines = ["boardsize 19", "clear_board", "komi 7.5", f"time_settings 0 {seconds} 1"]
    for m in range(moves):
        lines.append("genmove " + ("b" if m % 2 == 0 else "w"))
    lines.append("quit")
    return "\n".join(lines) + "\n"


def run(cmd, script, env=None, timeout=1200):
    t0 = time.time()
    r = subprocess.run(cmd, input=script, capture_output=True, text=True, env=env, timeout=timeout)
    text = r.stdout + r.stderr
    per_move = [dict(visits=int(a), nodes=int(b), playouts=int(c), playouts_per_s=int(d)) for a, b, c, d in STATS.findall(text)]
    moves = re.findall(r"^= ([A-T]\d+|pass|resign)\s*$", r.stdout, flags=re.M | re.I)
    return r.returncode, text, per_move, moves, time.time() - t0


def summarize(name, rc, text, per_move, moves, wall, extra):
    out = {"engine": name, "rc": rc, "moves": moves, "per_move": per_move, "wall_s": round(wall, 1)}
    nb = re.findall(r"(\d+) (predictions|evaluations) in\s+([\d.]+) seconds -> (\d+) p/s", text)
    if nb:
        out["netbench"] = {what: {"n": int(n), "seconds": float(sec), "per_s": int(ps)} for n, what, sec, ps in nb}
    if per_move:
        out["playouts_per_s_mean"] = sum(m["playouts_per_s"] for m in per_move) / len(per_move)
        out["playouts_total"] = sum(m["playouts"] for m in per_move)
    m = BATCH.search(text)
    if m:
        out["nn_positions"] = int(m.group(1)); out["device_batches"] = int(m.group(2))
        out["mean_device_batch"] = float(m.group(3)); out["nn_requests"] = int(m.group(4))
    fp = re.search(r"feature planes: ([\d.]+) us per position", text)
    if fp:
        out["feature_planes_us_per_position"] = float(fp.group(1))
    out.update(extra)
    if rc != 0 or (not per_move and not nb):
        out["tail"] = text[-1500:]
    return out


def parser():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=int, default=5)
    ap.add_argument("--moves", type=int, default=4)
    ap.add_argument("--threads", type=int, default=16)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--max-outstanding", type=int, default=2)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--netbench", action="store_true", help="run the GTP `netbench` command (Network::benchmark) instead of genmove")
    ap.add_argument("--extra", default="", help="extra engine options, e.g. '--mature_threshold 2 --eval_thresh 0'")
    return ap


def script_for(args) -> str:
    return gtp_script(args.seconds, args.moves) if not args.netbench else "boardsize 19\nclear_board\nnetbench\nquit\n"


def run_ours(args) -> dict:
    if not os.path.exists(WEIGHTS):
        sys.path.insert(0, ROOT)
        from engine import build as eb
        eb.write_synth_weights()
    cmd = [OURS, "-g", "-t", str(args.threads), "--noponder", "--nobook", "--lagbuffer", "0", "--weights", WEIGHTS,
           "--max-outstanding", str(args.max_outstanding)]
    if args.batch:
        cmd += ["--batch", str(args.batch)]
    for g in range(args.gpus):
        cmd += ["--gpu", str(g)]
    cmd += args.extra.split()
    return summarize("leela_b200_engine", *run(cmd, script_for(args)),
                     {"threads": args.threads, "gpus": args.gpus, "think_s": args.seconds,
                      "max_outstanding": args.max_outstanding, "extra": args.extra})


def main():
    print(json.dumps(run_ours(parser().parse_args())), flush=True)


if __name__ == "__main__":
    main()
