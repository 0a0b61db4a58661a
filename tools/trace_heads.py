"""Timeline of the heads launch relative to the end of the trunk launch (option "trace"): per block
%globaltimer stamps at block start (ev0) and end (ev8). Run on a B200."""
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
for generic in (1,):
    for _ in range(20):
        ev.eval_both_device(*a, stream=st.cuda_stream)
    torch.cuda.synchronize()
    ev.set_option("trace", 1)
    ev.eval_both_device(*a, stream=st.cuda_stream)
    torch.cuda.synchronize()
    ev.read_trace()
    for rep in range(2):
        for _ in range(3):
            ev.eval_both_device(*a, stream=st.cuda_stream)
        torch.cuda.synchronize()
        tr = ev.read_trace().astype(np.int64)
        trunk_start = tr[:, 95, 0].min(); trunk_end = tr[:, 95, 2].max()
        h = tr[:, 94, :9] - trunk_end
        nv = 144   # value blocks: 16 groups of 16 positions x 9 slices of the inner-product matrix (the first 148 blocks are traced)
        val, pol = h[:nv], h[nv:]
        print(f"generic={generic} rep {rep}: trunk {1e-3 * (trunk_end - trunk_start):.1f} us; value blocks ({nv}) stamps rel. to trunk end, us:")
        for e in range(9):
            col = val[:, e][tr[:nv, 94, e] > 0]
            if len(col):
                print(f"   ev{e}: min {col.min() / 1e3:6.2f} median {np.median(col) / 1e3:6.2f} max {col.max() / 1e3:6.2f}")
        for e in (0, 8):
            col = pol[:, e][tr[nv:, 94, e] > 0]
            if len(col):
                print(f"   policy blocks ev{e}: min {col.min() / 1e3:6.2f} median {np.median(col) / 1e3:6.2f} max {col.max() / 1e3:6.2f}")
    ev.set_option("trace", 0)
