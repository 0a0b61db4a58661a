"""Trunk / step time at batch 256 in plain and in precise (split-operand) mode, back-to-back steps on one box."""
import os, sys
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
a = (pp.data_ptr(), vp.data_ptr(), rot.data_ptr(), B, 0.75, probs.data_ptr(), win.data_ptr())
for precise in (0, 1, 0, 1):
    ev.set_option("precise", precise)
    for _ in range(10):
        ev.eval_both_device(*a, stream=st.cuda_stream)
    torch.cuda.synchronize()
    ev.set_option("profile_trunk", 1); ev.get_option("trunk_ns")
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(100):
        ev.eval_both_device(*a, stream=st.cuda_stream)
    e1.record(st); torch.cuda.synchronize()
    step = e0.elapsed_time(e1) / 100 * 1e3
    print(f"precise={precise} B={B}: step {step:8.1f} us, trunk {ev.get_option('trunk_ns')/100/1e3:8.1f} us, "
          f"{B / step * 1e6:,.0f} positions/s", flush=True)
    ev.set_option("profile_trunk", 0)
