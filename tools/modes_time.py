import sys, json, time, numpy as np, torch
sys.path.insert(0, '.')
from leela_b200 import capi, synth
sys.path.insert(0, 'tests')
import parity_report
ev = capi.Evaluator(policy=synth.policy_weights(), value=synth.value_weights())
b = np.load('tests/golden/bench_positions.npz')
dev = torch.device('cuda', 0)
B = 256
pp = torch.from_numpy(b['policy_planes'][:B].astype(np.int32)).to(dev); vp = torch.from_numpy(b['value_planes'][:B].astype(np.int32)).to(dev)
rot = torch.from_numpy(b['rotation'][:B].copy()).to(dev)
probs = torch.empty((B, 361), dtype=torch.float32, device=dev); win = torch.empty((B,), dtype=torch.float32, device=dev)
st = torch.cuda.Stream(dev); torch.cuda.set_stream(st)
out = []
for mode in [(0, 0), (0, 1), (1, 1), (0, 2), (2, 2)]:
    r = parity_report.report(ev, mode=mode)
    ev.set_precision(*mode)
    for i in range(10):
        ev.eval_both_device(pp.data_ptr(), vp.data_ptr(), rot.data_ptr(), B, 0.75, probs.data_ptr(), win.data_ptr(), stream=st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for i in range(100):
        ev.eval_both_device(pp.data_ptr(), vp.data_ptr(), rot.data_ptr(), B, 0.75, probs.data_ptr(), win.data_ptr(), stream=st.cuda_stream)
    e1.record(st); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 10
    r['step_us_100_back_to_back'] = us; r['positions_per_s'] = B / us * 1e6
    print(mode, 'policy max %.2e value max %.2e top1 %d/%d/%d  step %.1f us  %.0f pos/s' % (r['policy_max_abs_err'], r['value_max_abs_err'], r['top1_agree'], r['top1_near_tie'], r['top1_disagree'], us, B / us * 1e6), flush=True)
    out.append(r)
json.dump(out, open('gpurun_out/r2_modes.json', 'w'), indent=1)
