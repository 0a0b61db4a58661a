"""Per-job means from a trunk timeline trace (option "trace"): MMA issue time, waits, epilogue time per work item.
python tools/trace_items.py [both|value|policy] [B]   (on a B200)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from leela_b200 import capi, synth

which = sys.argv[1] if len(sys.argv) > 1 else "both"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
g = np.load(os.path.join(ROOT, "tests", "golden", "bench_positions.npz"))
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(dev); torch.cuda.set_stream(st)
pp = torch.from_numpy(g["policy_planes"][:B].astype(np.int32)).to(dev)
vp = torch.from_numpy(g["value_planes"][:B].astype(np.int32)).to(dev)
rot = torch.from_numpy(g["rotation"][:B].copy()).to(dev)
probs = torch.empty((B, 361), device=dev); win = torch.empty((B,), device=dev)
ev = capi.Evaluator(policy=synth.policy_weights(), value=synth.value_weights())
ev.set_option("max_batch", 512)
for kv in filter(None, os.environ.get("LB2_OPTS", "").split(",")):   # e.g. LB2_OPTS=resident_weights=1,policy_clusters=48
    ev.set_option(kv.split("=")[0], int(kv.split("=")[1]))
a = (pp.data_ptr(), vp.data_ptr(), rot.data_ptr(), B, 0.75, probs.data_ptr() if which != "value" else None,
     win.data_ptr() if which != "policy" else None)
for _ in range(20):
    ev.eval_both_device(*a, stream=st.cuda_stream)
torch.cuda.synchronize()
ev.set_option("trace", 1)
ev.eval_both_device(*a, stream=st.cuda_stream); torch.cuda.synchronize(); ev.read_trace()
for _ in range(3):
    ev.eval_both_device(*a, stream=st.cuda_stream)
torch.cuda.synchronize()
T = ev.read_trace().astype(np.int64)
print("launch %.1f us" % ((T[:, 95, 2].max() - T[:, 95, 0][T[:, 95, 0] > 0].min()) / 1e3))
t0 = T[:, 95, 0][T[:, 95, 0] > 0].min()
lead = T[0::2]
busy = []
for c in range(lead.shape[0]):
    n = int((lead[c, :94, 7] > 0).sum())
    j = lead[c, :n, 15] >> 21
    busy.append(((lead[c, :n, 7] - lead[c, :n, 6]).sum() / 1e3, n, (lead[c, n - 1, 7] - t0) / 1e3 if n else 0, int((j[: n] == j[0]).sum()) if n else 0))
busy = np.array(busy)
print("per cluster: MMA issue time min %.0f median %.0f max %.0f us; items min %d max %d; last MMA issued at min %.0f median %.0f max %.0f us" % (
    busy[:, 0].min(), np.median(busy[:, 0]), busy[:, 0].max(), busy[:, 1].min(), busy[:, 1].max(), busy[:, 2].min(), np.median(busy[:, 2]), busy[:, 2].max()))
for rank in (0, 1):
    L = T[rank::2]
    rows = {}
    for c in range(L.shape[0]):
        n = int((L[c, :94, 10] > 0).sum())
        for i in range(n):
            j = int(L[c, i, 15] >> 21)
            r = rows.setdefault(j, [])
            r.append((L[c, i, 7] - L[c, i, 6], L[c, i, 6] - L[c, i, 5], L[c, i, 5] - L[c, i, 4], L[c, i, 10] - L[c, i, 9], L[c, i, 9] - L[c, i, 8]))
    print("rank %d: job  items  mma-issue  wait-1st-stage  wait-acc  epilogue(tfull->drained)  epilogue-wait-tfull   (us, warp 2 = quadrant 2 for the epilogue)" % rank)
    for j in sorted(rows):
        m = np.array(rows[j]).mean(0) / 1e3
        print("   %4d %6d %9.2f %12.2f %10.2f %14.2f %18.2f" % (j, len(rows[j]), *m))
