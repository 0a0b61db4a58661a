"""Load balance / utilisation over time from a saved trace (gpurun_out/trace_<which>_<B>.npy)."""
import sys
import numpy as np
T = np.load(sys.argv[1]).astype(np.int64)
T = T.copy(); T[:, -1, :] = 0
lead = T[0::2] if (T[1::2, :, 7] > 0).sum() == 0 else T
t0 = lead[:, :, 0][lead[:, :, 0] > 0].min()
v = lead[:, :, 7] > 0
n = v.sum(1)
last_issue = np.array([lead[c, n[c] - 1, 7] for c in range(lead.shape[0])]) - t0
print('items per cluster: min %d max %d' % (n.min(), n.max()))
print('last MMA issued (us): min %.1f median %.1f max %.1f' % (last_issue.min() / 1e3, np.median(last_issue) / 1e3, last_issue.max() / 1e3))
issue = (lead[:, :, 7] - lead[:, :, 6]) * v
print('sum issue per cluster (us): mean %.1f min %.1f max %.1f' % (issue.sum(1).mean() / 1e3, issue.sum(1).min() / 1e3, issue.sum(1).max() / 1e3))
end = last_issue.max()
ts = np.linspace(0, end, 25)
out = []
for a, b in zip(ts[:-1], ts[1:]):
    busy = 0
    for c in range(lead.shape[0]):
        s = np.clip(lead[c, :n[c], 6] - t0, a, b); e = np.clip(lead[c, :n[c], 7] - t0, a, b)
        busy += (e - s).sum()
    out.append(busy / ((b - a) * lead.shape[0]))
print('issue-active fraction over time:', ' '.join('%.2f' % x for x in out))
