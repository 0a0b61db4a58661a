"""Every precision mode under every launch form (graphs, resident weights, per-layer launches, single CTA): value and policy error, run-to-run determinism, per-layer probe of the stored activations."""
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from leela_b200 import capi, synth
from oracle import oracle

g = np.load(os.path.join(ROOT, "tests/golden/bench_golden.npz")); b = np.load(os.path.join(ROOT, "tests/golden/bench_positions.npz"))
n = 256
pp, vp, rot = b["policy_planes"][:n], b["value_planes"][:n], b["rotation"][:n]
ev = capi.Evaluator(policy=synth.policy_weights(), value=synth.value_weights())
T = float(g["softmax_temp"])
def errs(tag):
    v = ev.eval_value(vp, rot); v2 = ev.eval_value(vp, rot)
    p = ev.eval_policy(pp, rot, T)
    print("%-44s value max %.2e (repeat diff %.1e)  policy max %.2e" % (tag, np.abs(v - g["value"][:n]).max(), np.abs(v - v2).max(), np.abs(p - g["policy"][:n]).max()), flush=True)
for mode in [(0, 1), (1, 1)]:
    ev.set_precision(*mode)
    for opts in [dict(), dict(use_graphs=0), dict(use_graphs=0, resident_weights=0), dict(use_graphs=0, trunk_mode=0), dict(use_graphs=0, cta_pair=0), dict(use_graphs=0, cta_pair=0, trunk_mode=0)]:
        for k, v in opts.items(): ev.set_option(k, v)
        errs(f"mode {mode} {opts}")
        for k in opts: ev.set_option(k, {"use_graphs": 1, "resident_weights": 2, "trunk_mode": 1, "cta_pair": 1}[k])
# per-layer probe, value net: how many stored activations differ from fp16(exact fp32 layer chain)?
vn = oracle.OracleNet(synth.value_weights())
for mode in [(0, 0), (0, 1), (0, 2)]:
    ev.set_precision(*mode)
    for L in (1, 2, 3, 6, 10):
        got = ev.debug_trunk(capi.VALUE, vp[:4], rot[:4], L, 64)
        worst, frac = 0.0, 0.0
        for i in range(4):
            want = oracle.trunk_activations(vn, vp[i], int(rot[i]))[L - 1]
            w16 = want.astype(np.float16).astype(np.float32)
            d = np.abs(got[i].reshape(want.shape) - w16)
            worst = max(worst, float((d / np.maximum(np.abs(want), 1e-2)).max())); frac += float((d > 0).mean()) / 4
        print(f"mode {mode} layer {L}: max rel err {worst:.2e}, fraction of entries != fp16(fp32 chain) {frac:.4f}", flush=True)
