"""Per-item timeline of the trunk kernel at batch B (debug). Prints where each role spends time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from leela_b200 import capi, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
which = sys.argv[2] if len(sys.argv) > 2 else "both"
g = np.load(os.path.join(ROOT, "tests", "golden", "bench_positions.npz"))
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(dev); torch.cuda.set_stream(st)
pp = torch.from_numpy(g["policy_planes"][:B].astype(np.int32)).to(dev)
vp = torch.from_numpy(g["value_planes"][:B].astype(np.int32)).to(dev)
rot = torch.from_numpy(g["rotation"][:B].copy()).to(dev)
probs = torch.empty((B, 361), device=dev); win = torch.empty((B,), device=dev)
ev = capi.Evaluator(policy=synth.policy_weights(), value=synth.value_weights())
ev.set_option("max_batch", max(B, 512))
if "LB2_PRECISION" in os.environ:   # e.g. LB2_PRECISION=0,0 (policy, value)
    ev.set_precision(*[int(x) for x in os.environ["LB2_PRECISION"].split(",")])
a = (pp.data_ptr(), vp.data_ptr(), rot.data_ptr(), B, 0.75, probs.data_ptr() if which != "value" else None,
     win.data_ptr() if which != "policy" else None)
for _ in range(5):
    ev.eval_both_device(*a, stream=st.cuda_stream)
torch.cuda.synchronize()
ev.set_option("trace", 1)
ev.eval_both_device(*a, stream=st.cuda_stream)
torch.cuda.synchronize()
T = ev.read_trace().astype(np.int64)          # [cta, item, event]
np.save(os.path.join(ROOT, "gpurun_out", f"trace_{which}_{B}.npy"), T)
clk = T[:, -1, :4]
mhz = (clk[:, 3] - clk[:, 1]) / np.maximum(clk[:, 2] - clk[:, 0], 1) * 1e3
print(f"effective SM clock during the kernel: median {np.median(mhz):.0f} MHz (min {mhz.min():.0f}, max {mhz.max():.0f})")
T = T.copy(); T[:, -1, :] = 0
valid = T[:, :, 7] > 0
t0 = T[:, :, 0][T[:, :, 0] > 0].min()
tend = T[:, :, 12].max()
print(f"{which} B={B}: kernel span {(tend - t0) / 1e3:.1f} us, items/cta {valid.sum(1).mean():.1f}")
def d(a, b):
    x = (T[:, :, b] - T[:, :, a])[valid & (T[:, :, a] > 0) & (T[:, :, b] > 0)]
    return f"mean {x.mean() / 1e3:7.2f} us  p50 {np.median(x) / 1e3:7.2f}  p90 {np.percentile(x, 90) / 1e3:7.2f}  max {x.max() / 1e3:7.2f}"
print("producer: dependency wait        (0->1) ", d(0, 1))
print("producer: first stage issued     (1->2) ", d(1, 2))
print("producer: all loads issued       (1->3) ", d(1, 3))
print("mma: wait accumulator drained    (4->5) ", d(4, 5))
print("mma: wait first stage data       (5->6) ", d(5, 6))
print("mma: issue all stages            (6->7) ", d(6, 7))
print("mma: item period (4 -> next 4)          ", end="")
x = np.diff(T[:, :, 4], axis=1)[valid[:, 1:] & valid[:, :-1]]
print(f"mean {x.mean() / 1e3:7.2f} us  p50 {np.median(x) / 1e3:7.2f}  p90 {np.percentile(x, 90) / 1e3:7.2f}")
print("epilogue: wait accumulator ready (8->9) ", d(8, 9))
print("epilogue: drain                  (9->10)", d(9, 10))
print("mma issued -> epilogue sees it   (7->9) ", d(7, 9))
print("publisher: drained -> stored     (10->11)", d(10, 11))
print("publisher: release               (11->12)", d(11, 12))
print("item latency: dep ready -> published (1->12)", d(1, 12))
