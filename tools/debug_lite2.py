"""Which e4m3 correction term is effective? (debug build with LB2_LITE_VARIANT)"""
import sys, os, numpy as np, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) == 1:
    for v in (0, 1, 2, 3):
        env = dict(os.environ, LB2_LIB=os.path.join(ROOT, "tools/_variants/liblb2_dbg.so"), LB2_LITE_VARIANT=str(v))
        subprocess.call([sys.executable, __file__, str(v)], env=env)
    sys.exit(0)
sys.path.insert(0, ROOT)
from leela_b200 import capi, synth
from oracle import oracle
b = np.load(os.path.join(ROOT, "tests/golden/bench_positions.npz")); g = np.load(os.path.join(ROOT, "tests/golden/bench_golden.npz"))
vp, rot = b["value_planes"][:64], b["rotation"][:64]
ev = capi.Evaluator(policy=synth.policy_weights(), value=synth.value_weights())
vn = oracle.OracleNet(synth.value_weights())
out = []
for L in (2, 3):
    got = ev.debug_trunk(capi.VALUE, vp[:4], rot[:4], L, 64)
    frac = 0.0
    for i in range(4):
        want = oracle.trunk_activations(vn, vp[i], int(rot[i]))[L - 1]
        frac += float((got[i].reshape(want.shape) != want.astype(np.float16).astype(np.float32)).mean()) / 4
    out.append("layer %d mismatch %.4f" % (L, frac))
v = ev.eval_value(vp, rot)
print("variant", sys.argv[1], "; ".join(out), "; value max err %.2e" % np.abs(v - g["value"][:64]).max(), flush=True)
