"""Parity report of the CUDA path against the reference over the correctness set
(tests/golden/bench_golden.npz: the reference's outputs for the 1024 bench positions, rotation i%8),
SURVEY.md section 8d: max / mean absolute error of the 361 probabilities and of the value, top-1 agreement
(ties within tolerance counted separately). Run on a B200: python tests/parity_report.py [out.json] [--precise]
(--precise: the split-operand mode, lb2_set_option("precise", 1))."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def report(ev=None, precise=False):
    from leela_b200 import capi, synth
    b = np.load(os.path.join(ROOT, "tests", "golden", "bench_positions.npz"))
    g = np.load(os.path.join(ROOT, "tests", "golden", "bench_golden.npz"))
    own = ev is None
    if own:
        ev = capi.Evaluator(policy=synth.policy_weights(), value=synth.value_weights())
    was = ev.get_option("precise")
    ev.set_option("precise", 1 if precise else 0)
    tol = 2e-4 if precise else 6e-3
    probs, win = ev.eval_both(b["policy_planes"], b["value_planes"], b["rotation"], float(g["softmax_temp"]))
    ev.set_option("precise", was)
    if own:
        ev.close()
    want_p, want_v = g["policy"].astype(np.float64), g["value"].astype(np.float64)
    dp, dv = np.abs(probs - want_p), np.abs(win - want_v)
    top_got, top_want = probs.argmax(1), want_p.argmax(1)
    same = top_got == top_want
    rows = np.arange(len(same))
    # a different top-1 counts as a tie when the reference itself rates the two moves within the tolerance
    near = ~same & (want_p[rows, top_want] - want_p[rows, top_got] < tol)
    return {
        "positions": int(len(same)), "policy_max_abs_err": float(dp.max()), "policy_mean_abs_err": float(dp.mean()),
        "policy_p99_9_abs_err": float(np.quantile(dp, 0.999)), "policy_max_err_per_position_median": float(np.median(dp.max(1))),
        "value_max_abs_err": float(dv.max()), "value_mean_abs_err": float(dv.mean()),
        "top1_agree": int(same.sum()), "top1_near_tie": int(near.sum()), "top1_disagree": int((~same & ~near).sum()),
        "frac_positions_within_1e-3": float((dp.max(1) < 1e-3).mean()), "frac_values_within_1e-3": float((dv < 1e-3).mean()),
        "operands": ("fp16 hi + fp16 lo activations and weights, three MMA terms (hi*Wh + hi*Wl + lo*Wh)" if precise
                     else "fp16 activations and weights") + ", fp32 accumulation / epilogue / heads",
        "tolerance_asserted": tol,
    }


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if a != "--precise"]
    r = report(precise="--precise" in sys.argv)
    print(json.dumps(r, indent=1))
    if args:
        json.dump(r, open(args[0], "w"), indent=1)
