"""Trunk / step time at batch 256 under option settings given as name=value[,name=value...] groups on the command line
(one group per measurement), back-to-back steps on one box: python tools/time_opts.py resident_weights=1 resident_weights=1,policy_clusters=50"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from leela_b200 import capi, synth

g = np.load(os.path.join(ROOT, "tests", "golden", "bench_positions.npz"))
B = 256
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(dev); torch.cuda.set_stream(st)
pp = torch.from_numpy(g["policy_planes"][:B].astype(np.int32)).to(dev)
vp = torch.from_numpy(g["value_planes"][:B].astype(np.int32)).to(dev)
rot = torch.from_numpy(g["rotation"][:B].copy()).to(dev)
probs = torch.empty((B, 361), device=dev); win = torch.empty((B,), device=dev)
ev = capi.Evaluator(policy=synth.policy_weights(), value=synth.value_weights())
ev.set_option("max_batch", 512)
a = (pp.data_ptr(), vp.data_ptr(), rot.data_ptr(), B, 0.75, probs.data_ptr(), win.data_ptr())
groups = ["baseline"] + sys.argv[1:] + ["baseline"]
defaults = {}
for grp in groups:
    for k, v in defaults.items():
        ev.set_option(k, v)
    if grp != "baseline":
        for kv in grp.split(","):
            k, v = kv.split("=")
            defaults.setdefault(k, ev.get_option(k))
            ev.set_option(k, int(v))
    for _ in range(10):
        ev.eval_both_device(*a, stream=st.cuda_stream)
    torch.cuda.synchronize()
    ev.set_option("profile_trunk", 1); ev.get_option("trunk_ns")
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(100):
        ev.eval_both_device(*a, stream=st.cuda_stream)
    e1.record(st); torch.cuda.synchronize()
    print(f"{grp:50s} step {e0.elapsed_time(e1) / 100 * 1e3:7.1f} us  trunk {ev.get_option('trunk_ns') / 100 / 1e3:7.1f} us  checksum {float(probs.sum()) + float(win.sum()):.6f}", flush=True)
    ev.set_option("profile_trunk", 0)
