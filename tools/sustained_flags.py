"""(Needs a debug build: python -c "from leela_b200 import build; build.build(force=True, defines=['LB2_DEBUG_KNOBS'], out='tools/_variants/liblb2_dbg.so')" and LB2_LIB=tools/_variants/liblb2_dbg.so — the product library compiles the timing knobs out.)
Sustained (seconds-long, power-capped) trunk time at batch 256 under LB2_DEBUG_FLAGS variants: how much of
the launch each component costs once the 1000 W cap, not the schedule, sets the pace. Results are wrong under
the flags; only the time matters. Usage (GPU box): python tools/sustained_flags.py [steps]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
CODE = f"""
import os, sys
sys.path.insert(0, {ROOT!r})
import numpy as np, torch
from leela_b200 import capi, synth
g = np.load(os.path.join({ROOT!r}, "tests", "golden", "bench_positions.npz"))
B = 256
dev = torch.device("cuda", 0); st = torch.cuda.Stream(dev); torch.cuda.set_stream(st)
pp = torch.from_numpy(g["policy_planes"][:B].astype(np.int32)).to(dev)
vp = torch.from_numpy(g["value_planes"][:B].astype(np.int32)).to(dev)
rot = torch.from_numpy(g["rotation"][:B].copy()).to(dev)
probs = torch.empty((B, 361), device=dev); win = torch.empty((B,), device=dev)
ev = capi.Evaluator(policy=synth.policy_weights(), value=synth.value_weights())
a = (pp.data_ptr(), vp.data_ptr(), rot.data_ptr(), B, 0.75, probs.data_ptr(), win.data_ptr())
for _ in range(20): ev.eval_both_device(*a, stream=st.cuda_stream)
torch.cuda.synchronize()
ev.set_option("profile_trunk", 1); ev.get_option("trunk_ns")
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(st)
for i in range({STEPS}):
    ev.eval_both_device(*a, stream=st.cuda_stream)
    if i % 500 == 499: torch.cuda.synchronize()
e1.record(st); torch.cuda.synchronize()
print("RES %.1f %.1f" % (e0.elapsed_time(e1) / {STEPS} * 1e3, ev.get_option("trunk_ns") / {STEPS} / 1e3))
"""
NAMES = {0: "baseline", 2: "all tap offsets 0 (aligned A reads)", 4: "epilogue without ELU math", 64: "no epilogue at all",
         8: "no B (weight) loads", 16: "no A (activation) loads", 24: "no operand loads", 32: "3x3 layers: first M half only (half the MMAs)"}
for f in (0, 2, 4, 64, 8, 16, 24, 32, 0):
    r = subprocess.run([sys.executable, "-c", CODE], capture_output=True, text=True, env=dict(os.environ, LB2_DEBUG_FLAGS=str(f)), timeout=600)
    line = [l for l in r.stdout.splitlines() if l.startswith("RES")]
    print(f"flags {f:3d} {NAMES[f]:48s} step/trunk us: {line[0][4:] if line else r.stderr[-300:]}", flush=True)
