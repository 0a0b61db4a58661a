"""Small run of every kernel mode (CTA pairs, single CTA, resident weights, lite / full precision, small-batch column splits,
whole layers at 70 positions, CUDA graphs, ensemble, 192-wide net) for
compute-sanitizer: `compute-sanitizer --tool memcheck|synccheck|initcheck python tests/sanitizer_run.py` (0 errors, round 1)."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from leela_b200 import capi, synth
g = np.load("tests/golden/ref_golden.npz")
ev = capi.Evaluator(policy=synth.policy_weights(), value=synth.value_weights())
ev.set_option("resident_weights", 0)
p, v = ev.eval_both(g["policy_planes"][:5], g["value_planes"][:5], g["rotation"][:5], 0.75)
print("both", float(np.abs(p - g["policy"][:5]).max()), float(np.abs(v - g["value"][:5]).max()))
pe, ve = ev.eval_ensemble(g["policy_planes"][:2], g["value_planes"][:2], 0.75)
print("ens", pe.sum(1), ve)
ev.set_option("resident_weights", 1)
p2, v2 = ev.eval_both(g["policy_planes"][:5], g["value_planes"][:5], g["rotation"][:5], 0.75)
print("resident identical", np.array_equal(p, p2), np.array_equal(v, v2))
ev.set_option("resident_weights", 0); ev.set_option("cta_pair", 0)
p3, v3 = ev.eval_both(g["policy_planes"][:5], g["value_planes"][:5], g["rotation"][:5], 0.75)
print("single identical", np.array_equal(p, p3), np.array_equal(v, v3))
for pair in (0, 1):
    ev.set_option("cta_pair", pair); ev.set_option("precise", 1)
    p4, v4 = ev.eval_both(g["policy_planes"][:5], g["value_planes"][:5], g["rotation"][:5], 0.75)
    print("precise pair=%d" % pair, float(np.abs(p4 - g["policy"][:5]).max()), float(np.abs(v4 - g["value"][:5]).max()))
ev.set_option("cta_pair", 1); ev.set_option("resident_weights", 2)
for mode in ((0, 1), (1, 1)):   # lite precision: e4m3 correction MMAs, e4m3 planes from the epilogue
    ev.set_precision(*mode)
    for n, sb in ((5, 48), (5, 0), (70, 48)):   # column splits / whole layers at a small batch / resident-weights launch
        ev.set_option("small_batch", sb)
        for rep in range(3):   # the third call of a shape replays its CUDA graph
            p5, v5 = ev.eval_both(g["policy_planes"][:n], g["value_planes"][:n], g["rotation"][:n], 0.75)
        print("lite", mode, n, sb, float(np.abs(p5 - g["policy"][:n]).max()), float(np.abs(v5 - g["value"][:n]).max()))
ev.close()
e2 = capi.Evaluator(policy=synth.policy192_weights())
q = e2.eval_policy(g["policy_planes"][:3], g["rotation"][:3], 0.75)
print("192", q.sum(1))
e2.set_option("precise", 1)
q2 = e2.eval_policy(g["policy_planes"][:3], g["rotation"][:3], 0.75)
print("192 precise", q2.sum(1), float(np.abs(q - q2).max()))
e2.close()
