"""BASELINE.json configs[2]: policy+value batch-size sweep 1..4096 on one B200 — latency per call
and throughput, kernels only (inputs in HBM, CUDA events) and end to end (pinned host buffers)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from leela_b200 import capi, synth, netdefs

g = np.load(os.path.join(ROOT, "tests", "golden", "bench_positions.npz"))
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(dev); torch.cuda.set_stream(st)
ev = capi.Evaluator(policy=synth.policy_weights(), value=synth.value_weights())
rows = []
for B in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096):
    idx = np.arange(B) % g["policy_planes"].shape[0]
    pp_h = torch.from_numpy(g["policy_planes"][idx].astype(np.int32)).pin_memory()
    vp_h = torch.from_numpy(g["value_planes"][idx].astype(np.int32)).pin_memory()
    rot_h = torch.from_numpy(g["rotation"][idx].copy()).pin_memory()
    pp, vp, rot = pp_h.to(dev), vp_h.to(dev), rot_h.to(dev)
    probs = torch.empty((B, 361), device=dev); win = torch.empty((B,), device=dev)
    probs_h = torch.empty((B, 361)).pin_memory(); win_h = torch.empty((B,)).pin_memory()
    reps = max(5, min(200, 20000 // max(B, 16)))
    a = (pp.data_ptr(), vp.data_ptr(), rot.data_ptr(), B, 0.75, probs.data_ptr(), win.data_ptr())
    for _ in range(3):
        ev.eval_both_device(*a, stream=st.cuda_stream)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        ev.eval_both_device(*a, stream=st.cuda_stream)
    e1.record(st); torch.cuda.synchronize()
    k_ms = e0.elapsed_time(e1) / reps
    for _ in range(2):
        ev.eval_both_raw(pp_h.data_ptr(), vp_h.data_ptr(), rot_h.data_ptr(), B, 0.75, probs_h.data_ptr(), win_h.data_ptr())
    t0 = time.perf_counter()
    for _ in range(reps):
        ev.eval_both_raw(pp_h.data_ptr(), vp_h.data_ptr(), rot_h.data_ptr(), B, 0.75, probs_h.data_ptr(), win_h.data_ptr())
    e_ms = (time.perf_counter() - t0) / reps * 1e3
    flops = (netdefs.POLICY_FLOPS + netdefs.VALUE_FLOPS) * B
    rows.append({"batch": B, "kernel_ms": k_ms, "kernel_pos_per_s": B / k_ms * 1e3, "pct_bf16_burst_peak": 100 * flops / (k_ms * 1e-3) / 1669.7e12,
                 "e2e_ms": e_ms, "e2e_pos_per_s": B / e_ms * 1e3})
    print(json.dumps(rows[-1]), flush=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "batch_sweep.json"), "w"), indent=1)
