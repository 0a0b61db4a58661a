import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["LB2_LIB"] = os.path.join(ROOT, "tools/_variants/liblb2_dbg.so")
os.environ["LB2_DEBUG_FLAGS"] = "128"
sys.path.insert(0, ROOT)
from leela_b200 import capi, synth
b = np.load(os.path.join(ROOT, "tests/golden/bench_positions.npz"))
ev = capi.Evaluator(policy=synth.policy_weights(), value=synth.value_weights())
ev.set_option("use_graphs", 0)
got = ev.debug_trunk(capi.VALUE, b["value_planes"][:2], b["rotation"][:2], 3, 64)
print("done", got.shape)
