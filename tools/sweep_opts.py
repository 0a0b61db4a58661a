"""Quick option sweeps at batch 256 (default precision): cluster split between the nets, device pass size. B200 only."""
import sys, os, json, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from leela_b200 import capi, synth
ev = capi.Evaluator(policy=synth.policy_weights(), value=synth.value_weights())
b = np.load(os.path.join(ROOT, "tests/golden/bench_positions.npz"))
dev = torch.device("cuda", 0)
B = 256
sets = []
for s in range(64):
    idx = (np.arange(B) + 3 * s) % 1024
    sets.append((torch.from_numpy(b["policy_planes"][idx].astype(np.int32)).to(dev), torch.from_numpy(b["value_planes"][idx].astype(np.int32)).to(dev),
                 torch.from_numpy(b["rotation"][idx].copy()).to(dev)))
probs = torch.empty((B, 361), dtype=torch.float32, device=dev); win = torch.empty((B,), dtype=torch.float32, device=dev)
st = torch.cuda.Stream(dev); torch.cuda.set_stream(st)
def run(n):
    for i in range(n):
        pp, vp, rot = sets[i % len(sets)]
        ev.eval_both_device(pp.data_ptr(), vp.data_ptr(), rot.data_ptr(), B, 0.75, probs.data_ptr(), win.data_ptr(), stream=st.cuda_stream)
def timed(tag, steps=60):
    run(10); torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); run(steps); e1.record(st); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / steps)
        import time; time.sleep(0.3)
    print("%-40s %.1f us/step  %.0f pos/s" % (tag, best, B / best * 1e6), flush=True)
    return best
timed("default")
for gp in (0,):
    ev.set_option("group_positions", gp); timed(f"group_positions {gp}", steps=200)
ev.set_option("resident_weights", 0); timed("resident_weights 0", steps=200); ev.set_option("resident_weights", 2)
ev.set_precision(0, 0)
for gp in (0, 128):
    ev.set_option("group_positions", gp); timed(f"fp16/fp16 group_positions {gp}", steps=200)
