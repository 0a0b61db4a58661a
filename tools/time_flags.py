"""Timing experiments: trunk launch time at batch 256 under LB2_DEBUG_FLAGS variants
(bit1: all tap offsets 0 = aligned A reads; bit2: epilogue without math). Results are wrong
under the flags; only the time matters."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from leela_b200 import capi, synth

g = np.load(os.path.join(ROOT, "tests", "golden", "bench_positions.npz"))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(dev); torch.cuda.set_stream(st)
pp = torch.from_numpy(g["policy_planes"][:B].astype(np.int32)).to(dev)
vp = torch.from_numpy(g["value_planes"][:B].astype(np.int32)).to(dev)
rot = torch.from_numpy(g["rotation"][:B].copy()).to(dev)
probs = torch.empty((B, 361), device=dev); win = torch.empty((B,), device=dev)
ev = capi.Evaluator(policy=synth.policy_weights(), value=synth.value_weights())
ev.set_option("max_batch", max(B, 512))
for mode in (1, 0, 2):
    ev.set_option("trunk_mode", 1 if mode else 0)
    ev.set_option("cta_pair", 0 if mode == 2 else 1)
    for which in ("both", "policy", "value"):
        a = (pp.data_ptr(), vp.data_ptr(), rot.data_ptr(), B, 0.75, probs.data_ptr() if which != "value" else None,
             win.data_ptr() if which != "policy" else None)
        for _ in range(5):
            ev.eval_both_device(*a, stream=st.cuda_stream)
        torch.cuda.synchronize()
        ev.set_option("profile_trunk", 1); ev.get_option("trunk_ns")
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(20):
            ev.eval_both_device(*a, stream=st.cuda_stream)
        e1.record(st); torch.cuda.synchronize()
        print(f"flags={os.environ.get('LB2_DEBUG_FLAGS','0')} mode={mode}{'(single-cta)' if mode == 2 else ''} {which:6s} B={B}: step {e0.elapsed_time(e1)/20*1e3:8.1f} us, "
              f"trunk {ev.get_option('trunk_ns')/20/1e3:8.1f} us", flush=True)
        ev.set_option("profile_trunk", 2)
        for _ in range(20):
            ev.eval_both_device(*a, stream=st.cuda_stream)
        torch.cuda.synchronize()
        if which == "both" and mode == 1:
            print("   segments (us): expand %.1f  trunk %.1f  heads %.1f" % tuple(ev.get_option(f"seg{k}") / 1e3 for k in range(3)), flush=True); ev.get_option("seg3")
        else:
            ev.get_option("seg3")
        ev.set_option("profile_trunk", 0)
