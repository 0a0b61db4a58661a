"""Parity report of the CUDA path against the reference over the correctness set
(tests/golden/bench_golden.npz: the reference's outputs for the 1024 bench positions, rotation i%8),
SURVEY.md section 8d: max / mean absolute error of the 361 probabilities and of the value, top-1 agreement
(ties within tolerance counted separately). Run on a B200: python tests/parity_report.py [out.json] [--mode P,V | --precise | --all]
(--mode P,V: the policy / value net's trunk precision, 0 = fp16 operands, 1 = lite (fp16 + e4m3 correction terms), 2 = full
split-operand precision; default 0,1 — the library's default; --precise = 2,2; --all: every supported combination)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


NAMES = {0: "fp16 activations and weights",
         1: "fp16 + e4m3 correction terms (hi*Wh + [e4m3(a) | e4m3(lo*2^12)] * [e4m3(Wl) ; e4m3(W)]), two MMA terms",
         2: "fp16 hi + fp16 lo activations and weights, three MMA terms (hi*Wh + hi*Wl + lo*Wh)"}
# what the tests assert for a net in each precision mode (measured: see profiles/r2_parity_*.json)
TOL_POLICY = {0: 6e-3, 1: 3e-4, 2: 2e-4}
TOL_VALUE = {0: 6e-3, 1: 3e-4, 2: 2e-4}


def report(ev=None, precise=False, mode=None):
    if mode is None:
        mode = (2, 2) if precise else (0, 1)
    from leela_b200 import capi, synth
    b = np.load(os.path.join(ROOT, "tests", "golden", "bench_positions.npz"))
    g = np.load(os.path.join(ROOT, "tests", "golden", "bench_golden.npz"))
    own = ev is None
    if own:
        ev = capi.Evaluator(policy=synth.policy_weights(), value=synth.value_weights())
    was = (ev.get_option("policy_precision"), ev.get_option("value_precision"))
    ev.set_precision(*mode)
    tol = TOL_POLICY[mode[0]]
    probs, win = ev.eval_both(b["policy_planes"], b["value_planes"], b["rotation"], float(g["softmax_temp"]))
    ev.set_precision(*was)
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
        "mode": {"policy_precision": mode[0], "value_precision": mode[1]},
        "operands": {"policy": NAMES[mode[0]], "value": NAMES[mode[1]], "rest": "fp32 accumulation / epilogue / heads"},
        "tolerance_asserted": {"policy": TOL_POLICY[mode[0]], "value": TOL_VALUE[mode[1]]},
    }


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    modes = [(0, 1)]
    if "--precise" in sys.argv:
        modes = [(2, 2)]
    if "--mode" in sys.argv:
        modes = [tuple(int(x) for x in sys.argv[sys.argv.index("--mode") + 1].split(","))]
        args = [a for a in args if "," not in a]
    if "--all" in sys.argv:
        modes = [(0, 0), (0, 1), (1, 1), (0, 2), (2, 2), (1, 0), (2, 0)]
    out = [report(mode=m) for m in modes]
    r = out[0] if len(out) == 1 else out
    print(json.dumps(r, indent=1))
    if args:
        json.dump(r, open(args[0], "w"), indent=1)
