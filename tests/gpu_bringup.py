"""GPU bring-up checks, one step per subprocess so a hang in one does not hide the others.
Usage (on the GPU box): python tests/gpu_bringup.py            # runs every step with timeouts
                        python tests/gpu_bringup.py step NAME  # one step in-process
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def golden():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_golden.npz")))


def stats(name, got, want):
    d = np.abs(got.astype(np.float64) - want.astype(np.float64))
    i = np.unravel_index(d.argmax(), d.shape)
    print(f"  {name}: max|d|={d.max():.3e} at {i} (got {got[i]:.5f} want {want[i]:.5f}) mean|d|={d.mean():.3e} "
          f"frac(|d|>1e-2)={(d > 1e-2).mean():.4f} nan={np.isnan(got).sum()}", flush=True)
    return d


def layer_step(kind, n_layers, n, mode):
    from leela_b200 import capi, synth
    from oracle import oracle
    g = golden()
    w = synth.policy_weights() if kind == 0 else synth.value_weights()
    planes = (g["policy_planes"] if kind == 0 else g["value_planes"])[:n]
    rot = g["rotation"][:n]
    ev = capi.Evaluator(policy=w if kind == 0 else None, value=w if kind == 1 else None)
    ev.set_option("trunk_mode", mode & 1)
    ev.set_option("cta_pair", (mode >> 1) & 1)
    print("backend:", ev.backend, "mode", mode, flush=True)
    c_out = w.convs[n_layers - 1].c_out
    t = time.time()
    got = ev.debug_trunk(kind, planes, rot, n_layers, c_out)
    print(f"  ran in {time.time() - t:.3f}s", flush=True)
    onet = oracle.OracleNet(w)
    worst = 0
    for i in range(n):
        acts = oracle.trunk_activations(onet, planes[i], int(rot[i]), emulate=7)
        d = stats(f"pos{i} layer{n_layers}", got[i], acts[n_layers - 1])
        worst = max(worst, d.max())
        if i == 0 and d.max() > 5e-2:
            # where are the errors? per-channel / per-pixel profile helps decode layout bugs
            print("   per-channel max:", np.round(d.max(1)[:16], 3))
            print("   per-pixel max (first 2 rows):", np.round(d.max(0)[:38], 3))
            print("   got[0,:8] ", np.round(got[i][0, :8], 4), "\n   want[0,:8]", np.round(acts[n_layers - 1][0, :8], 4))
    print("RESULT", "OK" if worst < 2e-2 else "MISMATCH", flush=True)


def full_step(n, mode):
    from leela_b200 import capi, synth
    from oracle import oracle
    g = golden()
    pw, vw = synth.policy_weights(), synth.value_weights()
    ev = capi.Evaluator(policy=pw, value=vw)
    ev.set_option("trunk_mode", mode & 1)
    ev.set_option("cta_pair", (mode >> 1) & 1)
    pp, vp, rot = g["policy_planes"][:n], g["value_planes"][:n], g["rotation"][:n]
    t = time.time()
    probs, win = ev.eval_both(pp, vp, rot, float(g["softmax_temp"]))
    print(f"  eval_both n={n} mode={mode} in {time.time() - t:.3f}s launches={ev.launch_count}", flush=True)
    stats("policy vs reference fp32", probs, g["policy"][:n])
    stats("value  vs reference fp32", win, g["value"][:n])
    pe = oracle.policy_forward(oracle.OracleNet(pw), pp, rot, float(g["softmax_temp"]), emulate=3)
    ve = oracle.value_forward(oracle.OracleNet(vw), vp, rot, emulate=3)
    d1 = stats("policy vs oracle fp16-emulation", probs, pe)
    d2 = stats("value  vs oracle fp16-emulation", win, ve)
    print("  top1 agreement vs reference:", (probs.argmax(1) == g["policy"][:n].argmax(1)).mean())
    print("RESULT", "OK" if d1.max() < 1e-3 and d2.max() < 1e-3 else "MISMATCH", flush=True)


STEPS = {
    # mode bit0: 1 = persistent dataflow launch, bit1: 1 = CTA pairs (cta_group::2)
    "p1_pair_layered": lambda: layer_step(0, 1, 2, 2),
    "p2_pair_layered": lambda: layer_step(0, 2, 2, 2),
    "v2_pair_layered": lambda: layer_step(1, 2, 2, 2),
    "p12_pair_mega": lambda: layer_step(0, 12, 3, 3),
    "full_pair_mega": lambda: full_step(16, 3),
    "full_pair_mega96": lambda: full_step(96, 3),
    "full_single_mega96": lambda: full_step(96, 1),
}

if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "step":
        STEPS[sys.argv[2]]()
        sys.exit(0)
    names = sys.argv[1:] or list(STEPS)
    for name in names:
        print(f"=== {name}", flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "step", name], timeout=90,
                               capture_output=True, text=True)
            print(r.stdout[-4000:], flush=True)
            if r.returncode != 0:
                print("  exit", r.returncode, r.stderr[-3000:], flush=True)
        except subprocess.TimeoutExpired as e:
            print("  TIMEOUT (hang)", (e.stdout or b"")[-2000:], flush=True)
