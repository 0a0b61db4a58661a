"""Device pass time at small batch sizes with and without the two-way column split of every layer (option small_batch)."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from leela_b200 import capi, synth
g = np.load(os.path.join(ROOT, "tests", "golden", "bench_positions.npz"))
dev = torch.device("cuda", 0); st = torch.cuda.Stream(dev); torch.cuda.set_stream(st)
ev = capi.Evaluator(policy=synth.policy_weights(), value=synth.value_weights())
for which in ("both", "value", "policy"):
    for B in (1, 8, 16, 32, 48, 64, 96, 128, 192):
        pp = torch.from_numpy(g["policy_planes"][:B].astype(np.int32)).to(dev); vp = torch.from_numpy(g["value_planes"][:B].astype(np.int32)).to(dev)
        rot = torch.from_numpy(g["rotation"][:B].copy()).to(dev)
        probs = torch.empty((B, 361), device=dev); win = torch.empty((B,), device=dev)
        a = (pp.data_ptr(), vp.data_ptr(), rot.data_ptr(), B, 0.75, probs.data_ptr() if which != "value" else None, win.data_ptr() if which != "policy" else None)
        out = []
        ref = None
        for sb in (0, 1 << 20):
            ev.set_option("small_batch", sb)
            for _ in range(5): ev.eval_both_device(*a, stream=st.cuda_stream)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(50): ev.eval_both_device(*a, stream=st.cuda_stream)
            e1.record(st); torch.cuda.synchronize()
            out.append(e0.elapsed_time(e1) * 20)
            res = (probs.clone(), win.clone())
            if ref is None: ref = res
            else: same = (which == "value" or torch.equal(ref[0], res[0])) and (which == "policy" or torch.equal(ref[1], res[1]))
        print(f"{which:6s} B={B:4d}: whole layers {out[0]:7.1f} us   two column splits {out[1]:7.1f} us   bit-identical {same}", flush=True)
